// nvJPEG through dlopen: the compressed-ingest entry point (ctag_detect_batch_jpeg) is the only user, and a
// box without libnvjpeg must still be able to load the detection library, so nothing links against it.  Only the types
// of <nvjpeg.h> are used at compile time.
#pragma once
#include <dlfcn.h>
#include <nvjpeg.h>

namespace ctag {

struct NvjpegApi {
  void* lib = nullptr;
  decltype(&nvjpegCreateEx) CreateEx = nullptr;
  decltype(&nvjpegDestroy) Destroy = nullptr;
  decltype(&nvjpegJpegStateCreate) JpegStateCreate = nullptr;
  decltype(&nvjpegJpegStateDestroy) JpegStateDestroy = nullptr;
  decltype(&nvjpegGetImageInfo) GetImageInfo = nullptr;
  decltype(&nvjpegDecodeBatchedInitialize) DecodeBatchedInitialize = nullptr;
  decltype(&nvjpegDecodeBatched) DecodeBatched = nullptr;
  bool ok = false;
};

inline const NvjpegApi& nvjpeg_api() {
  static NvjpegApi api = [] {
    NvjpegApi a;
    for (const char* name : {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (a.lib) break;
    }
    if (!a.lib) return a;
#define CTAG_NVJ(sym) a.sym = reinterpret_cast<decltype(a.sym)>(dlsym(a.lib, "nvjpeg" #sym))
    CTAG_NVJ(CreateEx);
    CTAG_NVJ(Destroy);
    CTAG_NVJ(JpegStateCreate);
    CTAG_NVJ(JpegStateDestroy);
    CTAG_NVJ(GetImageInfo);
    CTAG_NVJ(DecodeBatchedInitialize);
    CTAG_NVJ(DecodeBatched);
#undef CTAG_NVJ
    a.ok = a.CreateEx && a.Destroy && a.JpegStateCreate && a.JpegStateDestroy && a.GetImageInfo && a.DecodeBatchedInitialize &&
           a.DecodeBatched;
    return a;
  }();
  return api;
}

}  // namespace ctag
