// Library-level entry points of libmla_b200.so: version, error strings, device probe.
#include <mutex>
#include "common.cuh"

namespace mla {

std::atomic<uint64_t> g_launches{0};

const DeviceInfo& device_info() {
  static DeviceInfo info{};
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { info.ok = MLA_E_NODEVICE; return; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { info.ok = MLA_E_NODEVICE; return; }
    info.device = dev;
    info.sm_count = prop.multiProcessorCount;
    info.coop = prop.cooperativeLaunch;
    info.smem_optin = (int)prop.sharedMemPerBlockOptin;
    info.cc_major = prop.major;
    // The fat binary holds sm_100a SASS only: anything else cannot run these kernels.
    info.ok = (prop.major == 10) ? 1 : MLA_E_NODEVICE;
  });
  return info;
}

}  // namespace mla

extern "C" int mla_abi_version(void) { return MLA_ABI_VERSION; }

extern "C" const char* mla_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case MLA_E_BADARG: return "mla: bad argument (null, misaligned or inconsistent pointers)";
    case MLA_E_SHAPE: return "mla: shape not supported by this kernel";
    case MLA_E_WORKSPACE: return "mla: workspace missing or too small";
    case MLA_E_NODEVICE: return "mla: no usable sm_100 device (or cooperative launch unsupported)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "mla: unknown error";
}

extern "C" int mla_device_sm_count(void) {
  const mla::DeviceInfo& di = mla::device_info();
  return di.ok == 1 ? di.sm_count : di.ok;
}

extern "C" uint64_t mla_launch_count(void) { return mla::g_launches.load(std::memory_order_relaxed); }
