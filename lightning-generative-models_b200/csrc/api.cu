#include <stdlib.h>
// C-ABI plumbing: error string, launch counter, capability probe, conv dispatch.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace b200dm {
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200DM_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}


static std::atomic<int> g_reserved_sms{0};
int reserved_sms() { return g_reserved_sms.load(std::memory_order_relaxed); }

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};  // process-wide: autograd runs backward on its own thread

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B200DM_ERR_CUDA;
  }
  return B200DM_OK;
}

int conv_fwd_simt(const b200dm_conv_desc* d, void* stream);
int conv_wgrad_simt(const b200dm_wgrad_desc* d, void* stream);
int conv_fwd_tc(const b200dm_conv_desc* d, void* stream);
int conv_wgrad_tc(const b200dm_wgrad_desc* d, void* stream);
int conv_gn_fwd_tc(const b200dm_conv_desc* d, const b200dm_gn_desc* gn, void* stream);
int conv_gn_supported_tc(const b200dm_conv_desc* d, const b200dm_gn_desc* gn);
bool tc_supported();

}  // namespace b200dm

using namespace b200dm;

extern "C" int b200dm_version(void) { return 100; }
extern "C" const char* b200dm_last_error(void) { return g_err; }
extern "C" int64_t b200dm_launch_count(void) { return g_launches.load(); }
extern "C" void b200dm_reset_launch_count(void) { g_launches = 0; }
extern "C" int b200dm_tc_available(void) { return tc_supported() ? 1 : 0; }
extern "C" int b200dm_set_reserved_sms(int32_t n) {
  B200DM_REQUIRE(n >= 0 && n <= 64, B200DM_ERR_SHAPE, "set_reserved_sms: n=%d out of range [0, 64]", n);
  g_reserved_sms.store(n);
  return B200DM_OK;
}

static int check_conv_common(int dtype, int mode, int ksize, int B, int H, int W, int Cin, int Cout) {
  B200DM_REQUIRE(dtype == B200DM_F32 || dtype == B200DM_BF16, B200DM_ERR_UNSUPPORTED, "conv: dtype %d", dtype);
  B200DM_REQUIRE(mode >= 0 && mode <= 3, B200DM_ERR_UNSUPPORTED, "conv: mode %d", mode);
  B200DM_REQUIRE(mode != 0 || ksize == 1 || ksize == 3, B200DM_ERR_UNSUPPORTED, "conv: ksize %d (1 or 3)", ksize);
  B200DM_REQUIRE(B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0, B200DM_ERR_SHAPE,
                 "conv: empty shape B=%d H=%d W=%d Cin=%d Cout=%d", B, H, W, Cin, Cout);
  return B200DM_OK;
}

extern "C" int b200dm_conv_fwd(const b200dm_conv_desc* d, void* stream) {
  B200DM_REQUIRE(d != nullptr, B200DM_ERR_SHAPE, "conv_fwd: null descriptor");
  int rc = check_conv_common(d->dtype, d->mode, d->ksize, d->B, d->H, d->W, d->Cin, d->Cout);
  if (rc) return rc;
  B200DM_REQUIRE(d->x && d->w && d->y, B200DM_ERR_SHAPE, "conv_fwd: null tensor pointer");
  B200DM_REQUIRE(d->x_ld >= d->Cin && d->y_ld >= d->Cout, B200DM_ERR_SHAPE, "conv_fwd: ld smaller than channel count");
  if (d->impl == 1) {
    B200DM_REQUIRE(d->dtype == B200DM_BF16, B200DM_ERR_UNSUPPORTED, "conv_fwd(tc): bf16 only");
    return conv_fwd_tc(d, stream);
  }
  B200DM_REQUIRE(d->impl == 0, B200DM_ERR_UNSUPPORTED, "conv_fwd: impl %d", d->impl);
  B200DM_REQUIRE(d->mode != 3, B200DM_ERR_UNSUPPORTED,
                 "conv_fwd: mode 3 (fused nearest-2x upsample + 3x3) is built for the tcgen05 path (impl 1) only");
  B200DM_REQUIRE(d->gn_part == nullptr, B200DM_ERR_UNSUPPORTED,
                 "conv_fwd: fused GroupNorm statistics are built for the tcgen05 path (impl 1) only");
  return conv_fwd_simt(d, stream);
}

extern "C" int b200dm_conv_gn_supported(const b200dm_conv_desc* d, const b200dm_gn_desc* gn) {
  if (!d || !gn || d->impl != 1 || d->gn_part) return 0;
  return conv_gn_supported_tc(d, gn);
}

extern "C" int b200dm_conv_gn_fwd(const b200dm_conv_desc* d, const b200dm_gn_desc* gn, void* stream) {
  B200DM_REQUIRE(d != nullptr && gn != nullptr, B200DM_ERR_SHAPE, "conv_gn_fwd: null descriptor");
  int rc = check_conv_common(d->dtype, d->mode, d->ksize, d->B, d->H, d->W, d->Cin, d->Cout);
  if (rc) return rc;
  B200DM_REQUIRE(d->x && d->w && d->y, B200DM_ERR_SHAPE, "conv_gn_fwd: null tensor pointer");
  B200DM_REQUIRE(d->x_ld >= d->Cin && d->y_ld >= d->Cout, B200DM_ERR_SHAPE, "conv_gn_fwd: ld smaller than channel count");
  B200DM_REQUIRE(d->impl == 1 && d->dtype == B200DM_BF16 && !d->gn_part && !d->accumulate, B200DM_ERR_UNSUPPORTED,
                 "conv_gn_fwd: tcgen05 path (impl 1, bf16) only, without gn_part / accumulate");
  return conv_gn_fwd_tc(d, gn, stream);
}

extern "C" int b200dm_conv_wgrad(const b200dm_wgrad_desc* d, void* stream) {
  B200DM_REQUIRE(d != nullptr, B200DM_ERR_SHAPE, "conv_wgrad: null descriptor");
  int rc = check_conv_common(d->dtype, d->mode, d->ksize, d->B, d->H, d->W, d->Cin, d->Cout);
  if (rc) return rc;
  B200DM_REQUIRE(d->mode != 2 && d->mode != 3, B200DM_ERR_UNSUPPORTED,
                 "conv_wgrad: mode %d has no weight gradient (use mode 1 for the down/up-shuffle pair)", d->mode);
  B200DM_REQUIRE(d->x && d->dy && d->dw, B200DM_ERR_SHAPE, "conv_wgrad: null tensor pointer");
  if (!d->accumulate) {
    int taps = d->mode == 0 ? d->ksize * d->ksize : 4;
    rc = b200dm_fill_f32(d->dw, (int64_t)taps * d->Cout * d->Cin, 0.f, stream);
    if (rc) return rc;
  }
  if (d->impl == 1) {
    B200DM_REQUIRE(d->dtype == B200DM_BF16, B200DM_ERR_UNSUPPORTED, "conv_wgrad(tc): bf16 only");
    return conv_wgrad_tc(d, stream);
  }
  B200DM_REQUIRE(d->impl == 0, B200DM_ERR_UNSUPPORTED, "conv_wgrad: impl %d", d->impl);
  return conv_wgrad_simt(d, stream);
}
