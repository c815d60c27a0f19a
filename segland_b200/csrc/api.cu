// Library-level entry points: ABI version, error strings, device check.
#include "common.cuh"

namespace sl {
static Env g_env;
static bool g_env_loaded = false;
static void load_env() {
  auto geti = [](const char* name, int dflt) {
    const char* v = getenv(name);
    return v != nullptr && *v != '\0' ? atoi(v) : dflt;
  };
  Env e;
  e.tc_pair = geti("SL_TC_PAIR", -1);
  e.tc_small = geti("SL_TC_SMALL", -1);
  e.tc_debug = geti("SL_TC_DEBUG", 0);
  e.prep_split = geti("SL_PREP_SPLIT", 0);
  e.post_fused_cm = geti("SL_POST_FUSED_CM", -1);
  e.post_prune = geti("SL_POST_PRUNE", 0);
  e.post_regs = geti("SL_POST_REGS", 1);
  e.tail_fused = geti("SL_TAIL_FUSED", 1);
  e.fg_mma = geti("SL_FG_MMA", 1);
  const char* d = getenv("SL_SMALL_DBG");
  e.small_dbg = d != nullptr ? static_cast<long long>(strtoull(d, nullptr, 10)) : 0;
  g_env = e;
  g_env_loaded = true;
}
const Env& env() {
  if (!g_env_loaded) load_env();          // benign race: every thread computes the same values
  return g_env;
}
}  // namespace sl

extern "C" int sl_env_reload(void) { sl::load_env(); return SL_OK; }

extern "C" int sl_abi_version(void) { return SL_ABI_VERSION; }

extern "C" const char* sl_error_string(int code) {
  switch (code) {
    case SL_OK: return "ok";
    case SL_EINVAL: return "segland_b200: size/shape argument out of the supported range";
    case SL_ENULL: return "segland_b200: required pointer is NULL";
    case SL_EALIGN: return "segland_b200: pointer or extent violates the required alignment";
    case SL_EUNSUPPORTED: return "segland_b200: device is not sm_100 (B200); no other target is built";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "segland_b200: unknown error code";
}

extern "C" int sl_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  return major == 10 ? SL_OK : SL_EUNSUPPORTED;
}
