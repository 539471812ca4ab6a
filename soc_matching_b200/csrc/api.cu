// api.cu -- bookkeeping entry points of the C ABI (include/socm_b200.h) and shared host helpers.
#include <stdarg.h>
#include <atomic>
#include <stdlib.h>

#include "common.cuh"

namespace socm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int validate_setting(const socm_setting* st) {
  SOCM_CHECK_ARG(st != nullptr, "setting is NULL");
  SOCM_CHECK_ARG(st->d >= 1 && st->d <= SOCM_MAX_DIM, "d=%d outside [1,%d]", st->d, SOCM_MAX_DIM);
  SOCM_CHECK_ARG(st->lmbd > 0.f, "lmbd must be positive");
  SOCM_CHECK_ARG(st->sigma && st->sigma_inv, "sigma / sigma_inv missing");
  switch (st->kind) {
    case SOCM_OU_QUADRATIC:
      SOCM_CHECK_ARG(st->A && st->P && st->Q, "OU_quadratic needs A, P, Q");
      break;
    case SOCM_OU_LINEAR:
      SOCM_CHECK_ARG(st->A && st->omega, "OU_linear needs A, omega");
      break;
    case SOCM_DOUBLE_WELL:
      SOCM_CHECK_ARG(st->kappa && st->nu, "double_well needs kappa, nu");
      break;
    case SOCM_MOLECULAR_DYNAMICS:
      SOCM_CHECK_ARG(st->kappa, "molecular_dynamics needs kappa");
      break;
    default:
      set_error("unknown setting kind %d (no fallback by design)", st->kind);
      return SOCM_ERR_UNSUPPORTED;
  }
  return SOCM_OK;
}

int validate_unet(const socm_unet* net, int d) {
  SOCM_CHECK_ARG(net != nullptr, "unet is NULL");
  SOCM_CHECK_ARG(net->d == d, "unet.d=%d != setting.d=%d", net->d, d);
  SOCM_CHECK_ARG(net->h0 >= 1 && net->h1 >= 1 && net->h2 >= 1, "bad hidden sizes");
  SOCM_CHECK_ARG(net->h0 <= 1024 && net->h1 <= 1024 && net->h2 <= 1024, "hidden sizes > 1024 unsupported");
  for (int l = 0; l < 9; ++l) SOCM_CHECK_ARG(net->w[l] && net->b[l], "unet layer %d has NULL weights", l);
  return SOCM_OK;
}

bool is_default_arch(const socm_unet* net) { return net->h0 == 256 && net->h1 == 128 && net->h2 == 64; }

static std::atomic<int> g_engine_override{-1};

int f16_default() {
  const int o = g_engine_override.load(std::memory_order_relaxed);
  if (o >= 0) return o;
  static const int v = [] {
    const char* e = getenv("SOCM_F16");
    return e == nullptr ? -1 : (e[0] == '1' ? 1 : 0);
  }();
  return v;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    cached_dev = dev;
  }
  return cached > 0 ? cached : 148;
}

}  // namespace socm

extern "C" {

int socm_abi_version(void) { return 1; }

const char* socm_last_error(void) { return socm::g_err; }

int socm_device_info(int* sm_count, int* smem_optin_bytes) {
  int dev = 0;
  SOCM_CUDA(cudaGetDevice(&dev));
  if (sm_count) SOCM_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (smem_optin_bytes)
    SOCM_CUDA(cudaDeviceGetAttribute(smem_optin_bytes, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  return SOCM_OK;
}

int64_t socm_unet_param_count(const socm_unet* net) {
  if (!net) return -1;
  const int64_t d = net->d, h0 = net->h0, h1 = net->h1, h2 = net->h2;
  return (d + 1) * h0 + h0 + h0 * h1 + h1 + h1 * h2 + h2 + (d + 1) * d + d + h0 * h0 + h0 + h1 * h1 + h1 +
         h2 * h1 + h1 + h1 * h0 + h0 + h0 * d + d;
}

}  // extern "C"

extern "C" int socm_set_default_engine(int32_t engine) {
  SOCM_CHECK_ARG(engine >= -1 && engine <= 1, "engine must be -1 (environment / default), 0 (3xTF32) or 1 (fp16 split)");
  socm::g_engine_override.store(engine, std::memory_order_relaxed);
  return SOCM_OK;
}
