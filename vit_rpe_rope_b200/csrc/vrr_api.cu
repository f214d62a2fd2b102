// extern "C" entry points of libvrr_b200.so (see include/vrr.h): argument validation, kernel
// family selection, workspace carving.  No allocation, no synchronisation, no exceptions.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"

namespace vrr {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_impl{VRR_IMPL_AUTO};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Per-DEVICE state (a process may drive several GPUs): 0 unknown, 1 ok, -1 not sm_100.  Plain
// atomics: racing first calls on one device compute the same values.
constexpr int kMaxDevices = 64;
static std::atomic<int> g_dev_state[kMaxDevices];
static std::atomic<int> g_dev_sms[kMaxDevices];
std::atomic<uint64_t> g_family_launches[3];  // [VRR_IMPL_SIMT], [VRR_IMPL_TCGEN05] dispatch counters
// 4: without bias the four-CTA kernel of attn_tc.cu at every N, with bias as 3;  3: whole-sequence kernel for N <= 256
// (attn_fwd_ws.cu), else attn_tc.cu;  2: attn_tc.cu always
static std::atomic<int> g_attn_fwd_variant{4};
static std::atomic<int> g_attn_bwd_variant{3};  // 3: attn_bwd_ws.cu for N <= 208 without bias, else attn_bwd_tc2.cu; 2: attn_bwd_tc2.cu always

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return dev;
}

int require_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return VRR_ERR_NO_DEVICE;
  }
  if (dev < 0 || dev >= kMaxDevices) {
    set_error("device ordinal %d out of range", dev);
    return VRR_ERR_NO_DEVICE;
  }
  // Driver-API calls made by the launchers (cuTensorMapEncodeTiled) need the primary context CURRENT in the
  // calling thread.  A runtime call that touches the device binds it; cudaGetDevice does not.  PyTorch's autograd
  // worker threads reach this library without one when their allocations are served from the caching allocator
  // (observed: CUDA_ERROR_INVALID_CONTEXT from the first backward of a process) - bind once per thread.
  static thread_local int bound_dev = -1;
  if (bound_dev != dev) {
    e = cudaFree(nullptr);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("cannot initialise the CUDA context of device %d: %s", dev, cudaGetErrorString(e));
      return VRR_ERR_NO_DEVICE;
    }
    bound_dev = dev;
  }
  const int state = g_dev_state[dev].load(std::memory_order_acquire);
  if (state == 1) return VRR_OK;
  int major = 0, minor = 0, sms = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    return VRR_ERR_NO_DEVICE;
  }
  if (major != 10) {
    set_error("device %d is sm_%d%d; libvrr_b200 is built for sm_100a only (no fallback path)", dev, major, minor);
    g_dev_state[dev].store(-1, std::memory_order_release);
    return VRR_ERR_NO_DEVICE;
  }
  g_dev_sms[dev].store(sms, std::memory_order_relaxed);
  g_dev_state[dev].store(1, std::memory_order_release);
  return VRR_OK;
}
int sm_count() {
  const int dev = current_device();
  const int n = (dev >= 0 && dev < kMaxDevices) ? g_dev_sms[dev].load(std::memory_order_relaxed) : 0;
  return n > 0 ? n : 148;
}

// One-time per (kernel, device) attribute set-up: `flags` is a function-local static array.
bool first_use_on_device(std::atomic<unsigned char>* flags) {
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return true;
  return flags[dev].exchange(1, std::memory_order_acq_rel) == 0;
}

static bool dtype_ok(int d) { return d == VRR_F32 || d == VRR_BF16; }
static bool dh_ok(int dh) { return dh == 16 || dh == 32 || dh == 64; }

static int check_bias(const vrr_bias_desc* b, int H, int N) {
  if (!b || b->mode == VRR_BIAS_NONE) return VRR_OK;
  VRR_REQUIRE(b->param != nullptr, VRR_ERR_INVALID_ARG, "bias: param is NULL");
  if (b->mode == VRR_BIAS_TABLE) {
    VRR_REQUIRE(b->heads == H && b->len == 2 * N - 1, VRR_ERR_INVALID_ARG,
                "bias table must be [H=%d][2N-1=%d], got [%d][%d] (the relative table is tied to the "
                "construction-time token count)", H, 2 * N - 1, b->heads, b->len);
  } else if (b->mode == VRR_BIAS_POLY) {
    VRR_REQUIRE((b->heads == 1 || b->heads == H) && b->len >= 1 && b->len <= 16, VRR_ERR_INVALID_ARG,
                "poly coefficients must be [1 or H][1..16], got [%d][%d]", b->heads, b->len);
    VRR_REQUIRE(b->grid >= 1 && b->grid * b->grid == N - 1, VRR_ERR_INVALID_ARG,
                "poly bias needs N-1 == grid^2 (N=%d, grid=%d)", N, b->grid);
  } else {
    set_error("bias: unknown mode %d", b->mode);
    return VRR_ERR_INVALID_ARG;
  }
  return VRR_OK;
}

}  // namespace vrr

using namespace vrr;

extern "C" {

int vrr_abi_version(void) { return VRR_ABI_VERSION; }
const char* vrr_last_error(void) { return g_err; }
int vrr_device_ok(void) { return require_device() == VRR_OK ? 1 : 0; }
int vrr_set_impl(int impl) {
  if (impl < VRR_IMPL_AUTO || impl > VRR_IMPL_TCGEN05) return g_impl.load();
  return g_impl.exchange(impl);
}
uint64_t vrr_launch_count(void) { return g_launches.load(); }
uint64_t vrr_family_count(int family) {
  if (family != VRR_IMPL_SIMT && family != VRR_IMPL_TCGEN05) return 0;
  return g_family_launches[family].load();
}
int vrr_debug_timestamps(void* device_buf) {  /* 64 x int64 device buffer, or NULL to switch off */
  attn_fwd_tc_set_debug((long long*)device_buf);
  attn_bwd_ws_set_debug((long long*)device_buf);
  return VRR_OK;
}
int vrr_set_option(const char* name, int value) {
  if (!name) return VRR_ERR_INVALID_ARG;
  if (!strcmp(name, "gemm_variant")) { gemm_tc_set_variant(value); return VRR_OK; }
  if (!strcmp(name, "simt_gemm_tile")) { gemm_simt_set_tile(value); return VRR_OK; }
  if (!strcmp(name, "ln_reg")) { layernorm_set_option(0, value); return VRR_OK; }
  if (!strcmp(name, "ln_bwd_minb")) { layernorm_set_option(1, value); return VRR_OK; }
  if (!strcmp(name, "attn_fwd_variant")) { g_attn_fwd_variant.store(value); return VRR_OK; }
  if (!strcmp(name, "attn_bwd_variant")) { g_attn_bwd_variant.store(value); return VRR_OK; }
  if (!strcmp(name, "attn_fwd_poly_exp")) { attn_fwd_tc_set_poly(value); return VRR_OK; }
  if (!strcmp(name, "attn_fwd_streams")) { attn_fwd_tc_set_streams(value); return VRR_OK; }
  if (!strcmp(name, "attn_fwd_table_bulk")) { attn_fwd_tc_set_table_bulk(value); return VRR_OK; }
  if (!strcmp(name, "attn_fwd_rescale_threshold_x100")) {
    attn_fwd_tc_set_threshold_x100(value);
    return VRR_OK;
  }
  set_error("vrr_set_option: unknown option '%s'", name);
  return VRR_ERR_INVALID_ARG;
}

size_t vrr_patch_embed_workspace_bytes(int B, int C, int Hi, int Wi, int P, int E, int dtype) {
  if (dtype != VRR_BF16 || P <= 0 || !patch_embed_tc_supported(B, C, Hi, Wi, P, E)) return 0;
  return patch_embed_tc_workspace_bytes(B, C, Hi, Wi, P);
}

int vrr_patch_embed_fwd(const void* images, const void* weight, const void* bias, const void* cls_token,
                        const void* pos_embed, void* tokens, void* workspace, size_t workspace_bytes, int B, int C,
                        int Hi, int Wi, int P, int E, int img_dtype, int dtype, int tok_dtype, void* stream) {
  VRR_REQUIRE(images && weight && bias && cls_token && tokens, VRR_ERR_INVALID_ARG, "patch_embed_fwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype) && dtype_ok(tok_dtype) && dtype_ok(img_dtype), VRR_ERR_INVALID_ARG,
              "patch_embed_fwd: bad dtype %d/%d/%d", img_dtype, dtype, tok_dtype);
  VRR_REQUIRE(B > 0 && C > 0 && P > 0 && E > 0 && Hi >= P && Wi >= P, VRR_ERR_INVALID_ARG,
              "patch_embed_fwd: bad sizes B=%d C=%d Hi=%d Wi=%d P=%d E=%d", B, C, Hi, Wi, P, E);
  if (int rc = require_device()) return rc;
  const int impl = g_impl.load();
  if (dtype == VRR_BF16 && impl != VRR_IMPL_SIMT && patch_embed_tc_supported(B, C, Hi, Wi, P, E)) {
    const size_t need = patch_embed_tc_workspace_bytes(B, C, Hi, Wi, P);
    VRR_REQUIRE(workspace && workspace_bytes >= need, VRR_ERR_WORKSPACE,
                "patch_embed_fwd: workspace %zu < %zu bytes (vrr_patch_embed_workspace_bytes)", workspace_bytes, need);
    VRR_REQUIRE(((uintptr_t)workspace & 255) == 0, VRR_ERR_INVALID_ARG, "patch_embed_fwd: workspace must be 256-byte aligned");
    VRR_COUNT_FAMILY(VRR_IMPL_TCGEN05);
    if (int rc = patch_embed_fwd_tc(images, weight, bias, pos_embed, tokens, workspace, B, C, Hi, Wi, P, E, img_dtype,
                                    tok_dtype, (cudaStream_t)stream))
      return rc;
    return patch_cls_rows(tokens, cls_token, B, (Hi / P) * (Wi / P), E, tok_dtype, (cudaStream_t)stream);
  }
  VRR_REQUIRE(impl != VRR_IMPL_TCGEN05 || dtype != VRR_BF16, VRR_ERR_UNSUPPORTED,
              "patch_embed_fwd: tcgen05 kernel forced but shape unsupported (P %% 8, C*P*P %% 64, E %% 64)");
  VRR_COUNT_FAMILY(VRR_IMPL_SIMT);
  return patch_embed_fwd_simt(images, weight, bias, cls_token, pos_embed, tokens, B, C, Hi, Wi, P, E, img_dtype, dtype,
                              tok_dtype, (cudaStream_t)stream);
}

int vrr_patch_unfold(const void* images, void* out, int B, int C, int Hi, int Wi, int P, int img_dtype, void* stream) {
  VRR_REQUIRE(images && out, VRR_ERR_INVALID_ARG, "patch_unfold: NULL pointer");
  VRR_REQUIRE(dtype_ok(img_dtype), VRR_ERR_INVALID_ARG, "patch_unfold: bad dtype %d", img_dtype);
  VRR_REQUIRE(B > 0 && C > 0 && P > 0 && Hi >= P && Wi >= P, VRR_ERR_INVALID_ARG, "patch_unfold: bad sizes");
  if (int rc = require_device()) return rc;
  return patch_unfold(images, out, B, C, Hi, Wi, P, img_dtype, (cudaStream_t)stream);
}

int vrr_patch_embed_bwd(const void* images, const void* d_tokens, void* d_weight, void* d_bias, void* d_cls,
                        void* d_pos, int B, int C, int Hi, int Wi, int P, int E, int img_dtype, int dtype,
                        int tok_dtype, void* stream) {
  VRR_REQUIRE(images && d_tokens && d_bias && d_cls, VRR_ERR_INVALID_ARG, "patch_embed_bwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype) && dtype_ok(tok_dtype) && dtype_ok(img_dtype), VRR_ERR_INVALID_ARG,
              "patch_embed_bwd: bad dtype %d/%d/%d", img_dtype, dtype, tok_dtype);
  VRR_REQUIRE(B > 0 && C > 0 && P > 0 && E > 0 && Hi >= P && Wi >= P, VRR_ERR_INVALID_ARG,
              "patch_embed_bwd: bad sizes");
  if (int rc = require_device()) return rc;
  return patch_embed_bwd_simt(images, d_tokens, (float*)d_weight, (float*)d_bias, (float*)d_cls, (float*)d_pos, B,
                              C, Hi, Wi, P, E, img_dtype, dtype, tok_dtype, (cudaStream_t)stream);
}

int vrr_qkv_rope_fwd(const void* x, const void* w_qkv, const float* cos_tab, const float* sin_tab, void* planes,
                     int B, int N, int E, int H, int rope_mode, int dtype, void* stream) {
  return vrr_qkv_rope_fwd_packed(x, w_qkv, cos_tab, sin_tab, nullptr, planes, B, N, E, H, rope_mode, dtype, stream);
}

int vrr_rope_pack_tables(const float* cos_tab, const float* sin_tab, float* packed, int heads, int rows, int half_dim,
                         void* stream) {
  VRR_REQUIRE(cos_tab && sin_tab && packed, VRR_ERR_INVALID_ARG, "rope_pack_tables: NULL pointer");
  VRR_REQUIRE(heads > 0 && rows > 0 && half_dim > 0, VRR_ERR_INVALID_ARG, "rope_pack_tables: bad sizes");
  if (int rc = require_device()) return rc;
  return rope_pack_tables(cos_tab, sin_tab, packed, heads, rows, half_dim, (cudaStream_t)stream);
}

int vrr_qkv_rope_fwd_packed(const void* x, const void* w_qkv, const float* cos_tab, const float* sin_tab,
                            const float* packed, void* planes, int B, int N, int E, int H, int rope_mode, int dtype,
                            void* stream) {
  VRR_REQUIRE(x && w_qkv && planes, VRR_ERR_INVALID_ARG, "qkv_rope_fwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "qkv_rope_fwd: bad dtype %d", dtype);
  VRR_REQUIRE(rope_mode >= VRR_ROPE_NONE && rope_mode <= VRR_ROPE_MIXED, VRR_ERR_INVALID_ARG,
              "qkv_rope_fwd: bad rope_mode %d", rope_mode);
  VRR_REQUIRE(rope_mode == VRR_ROPE_NONE || (cos_tab && sin_tab), VRR_ERR_INVALID_ARG,
              "qkv_rope_fwd: cos/sin required for rope_mode %d", rope_mode);
  VRR_REQUIRE(B > 0 && N > 0 && H > 0 && E > 0 && E % H == 0, VRR_ERR_INVALID_ARG, "qkv_rope_fwd: bad sizes");
  VRR_REQUIRE(dh_ok(E / H), VRR_ERR_UNSUPPORTED, "qkv_rope_fwd: head dim %d unsupported (16, 32, 64)", E / H);
  if (int rc = require_device()) return rc;
  const int impl = g_impl.load();
  if (dtype == VRR_BF16 && impl != VRR_IMPL_SIMT && qkv_rope_fwd_tc_supported(B, N, E, H)) {
    VRR_COUNT_FAMILY(VRR_IMPL_TCGEN05);
    if (gemm_tc_variant() == 2)
      return qkv_rope_fwd_tc2(x, w_qkv, cos_tab, sin_tab, packed, planes, B, N, E, H, rope_mode, (cudaStream_t)stream);
    return qkv_rope_fwd_tc(x, w_qkv, cos_tab, sin_tab, planes, B, N, E, H, rope_mode, (cudaStream_t)stream);
  }
  VRR_REQUIRE(impl != VRR_IMPL_TCGEN05, VRR_ERR_UNSUPPORTED,
              "qkv_rope_fwd: tcgen05 kernel forced but shape/dtype unsupported (bf16, Dh=64, E%%64==0)");
  VRR_COUNT_FAMILY(VRR_IMPL_SIMT);
  return qkv_rope_fwd_simt(x, w_qkv, cos_tab, sin_tab, planes, B, N, E, H, rope_mode, dtype, (cudaStream_t)stream);
}

int vrr_qkv_rope_bwd(const void* d_planes, const void* planes, const float* cos_tab, const float* sin_tab,
                     void* d_qkv, float* d_cos, float* d_sin, int B, int N, int E, int H, int rope_mode,
                     int dtype, void* stream) {
  VRR_REQUIRE(d_planes && planes && d_qkv, VRR_ERR_INVALID_ARG, "qkv_rope_bwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "qkv_rope_bwd: bad dtype %d", dtype);
  VRR_REQUIRE(rope_mode >= VRR_ROPE_NONE && rope_mode <= VRR_ROPE_MIXED, VRR_ERR_INVALID_ARG,
              "qkv_rope_bwd: bad rope_mode %d", rope_mode);
  VRR_REQUIRE(rope_mode == VRR_ROPE_NONE || (cos_tab && sin_tab), VRR_ERR_INVALID_ARG,
              "qkv_rope_bwd: cos/sin required");
  VRR_REQUIRE((d_cos == nullptr) == (d_sin == nullptr), VRR_ERR_INVALID_ARG,
              "qkv_rope_bwd: d_cos and d_sin must both be given or both NULL");
  VRR_REQUIRE(B > 0 && N > 0 && H > 0 && E > 0 && E % H == 0 && dh_ok(E / H), VRR_ERR_INVALID_ARG,
              "qkv_rope_bwd: bad sizes");
  if (int rc = require_device()) return rc;
  return qkv_rope_bwd(d_planes, planes, cos_tab, sin_tab, d_qkv, d_cos, d_sin, B, N, E, H, rope_mode, dtype,
                      (cudaStream_t)stream);
}

int vrr_rope_apply(const void* q_in, const void* k_in, const float* cos_tab, const float* sin_tab, void* q_out,
                   void* k_out, int B, int H, int Nr, int Dh, int rope_mode, int inverse, int dtype,
                   void* stream) {
  VRR_REQUIRE(q_in && k_in && cos_tab && sin_tab && q_out && k_out, VRR_ERR_INVALID_ARG, "rope_apply: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "rope_apply: bad dtype %d", dtype);
  VRR_REQUIRE(rope_mode == VRR_ROPE_AXIAL || rope_mode == VRR_ROPE_MIXED, VRR_ERR_INVALID_ARG,
              "rope_apply: bad rope_mode %d", rope_mode);
  VRR_REQUIRE(B > 0 && H > 0 && Nr > 0 && Dh > 0 && Dh % 2 == 0, VRR_ERR_INVALID_ARG, "rope_apply: bad sizes");
  if (int rc = require_device()) return rc;
  return rope_apply(q_in, k_in, cos_tab, sin_tab, q_out, k_out, B, H, Nr, Dh, rope_mode, inverse, dtype,
                    (cudaStream_t)stream);
}

int vrr_rope_table_grad(const void* q_in, const void* k_in, const void* dq, const void* dk, float* d_cos,
                        float* d_sin, int B, int H, int Nr, int Dh, int rope_mode, int dtype, void* stream) {
  VRR_REQUIRE(q_in && k_in && dq && dk && d_cos && d_sin, VRR_ERR_INVALID_ARG, "rope_table_grad: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "rope_table_grad: bad dtype %d", dtype);
  VRR_REQUIRE(rope_mode == VRR_ROPE_AXIAL || rope_mode == VRR_ROPE_MIXED, VRR_ERR_INVALID_ARG,
              "rope_table_grad: bad rope_mode %d", rope_mode);
  VRR_REQUIRE(B > 0 && H > 0 && Nr > 0 && Dh > 0 && Dh % 2 == 0, VRR_ERR_INVALID_ARG, "rope_table_grad: bad sizes");
  if (int rc = require_device()) return rc;
  return rope_table_grad(q_in, k_in, dq, dk, d_cos, d_sin, B, H, Nr, Dh, rope_mode, dtype, (cudaStream_t)stream);
}

int vrr_gemm(const void* a, const void* b, void* c, int M, int N, int K, int trans_a, int trans_b, int dtype,
             int c_dtype, void* stream) {
  VRR_REQUIRE(a && b && c, VRR_ERR_INVALID_ARG, "gemm: NULL pointer");
  VRR_REQUIRE(M > 0 && N > 0 && K > 0, VRR_ERR_INVALID_ARG, "gemm: bad sizes");
  VRR_REQUIRE(dtype_ok(dtype) && dtype_ok(c_dtype), VRR_ERR_INVALID_ARG, "gemm: bad dtype %d/%d", dtype, c_dtype);
  if (int rc = require_device()) return rc;
  return gemm_simt(a, b, c, M, N, K, trans_a, trans_b, dtype, c_dtype, (cudaStream_t)stream);
}

int vrr_gemm_ex(const void* a, const void* b, void* c, void* c2, const float* bias, int M, int N, int K, int trans_a,
                int trans_b, int dtype, int c_dtype, int epilogue, int accumulate, void* stream) {
  VRR_REQUIRE(a && b && c, VRR_ERR_INVALID_ARG, "gemm_ex: NULL pointer");
  VRR_REQUIRE(M > 0 && N > 0 && K > 0, VRR_ERR_INVALID_ARG, "gemm_ex: bad sizes");
  VRR_REQUIRE(dtype_ok(dtype) && dtype_ok(c_dtype), VRR_ERR_INVALID_ARG, "gemm_ex: bad dtype %d/%d", dtype, c_dtype);
  VRR_REQUIRE(epilogue >= VRR_EPI_NONE && epilogue <= VRR_EPI_BIAS_GELU_ACT, VRR_ERR_INVALID_ARG, "gemm_ex: bad epilogue %d",
              epilogue);
  VRR_REQUIRE(epilogue == VRR_EPI_NONE || epilogue == VRR_EPI_MUL || bias, VRR_ERR_INVALID_ARG,
              "gemm_ex: the bias epilogues need `bias`");
  VRR_REQUIRE(epilogue < VRR_EPI_BIAS_GELU || epilogue == VRR_EPI_BIAS_GELU_ACT || c2, VRR_ERR_INVALID_ARG,
              "gemm_ex: the two-output GELU epilogues and MUL need `c2`");
  if (int rc = require_device()) return rc;
  const int impl = g_impl.load();
  if (dtype == VRR_BF16 && impl != VRR_IMPL_SIMT && gemm_bf16_tc_supported(M, N, K, trans_a, trans_b, c_dtype, epilogue)) {
    VRR_COUNT_FAMILY(VRR_IMPL_TCGEN05);
    return gemm_bf16_tc(a, b, c, c2, bias, M, N, K, trans_a, trans_b, c_dtype, epilogue, accumulate, (cudaStream_t)stream);
  }
  VRR_REQUIRE(impl != VRR_IMPL_TCGEN05 || dtype != VRR_BF16, VRR_ERR_UNSUPPORTED,
              "gemm_ex: tcgen05 kernel forced but M=%d N=%d K=%d (trans %d/%d) needs N %% 8 == 0, K %% 8 == 0 "
              "(M %% 8 == 0 when trans_a)", M, N, K, trans_a, trans_b);
  VRR_REQUIRE(!accumulate, VRR_ERR_UNSUPPORTED, "gemm_ex: accumulate is implemented by the tcgen05 kernel only");
  VRR_COUNT_FAMILY(VRR_IMPL_SIMT);
  if (epilogue == VRR_EPI_MUL) {
    VRR_REQUIRE(c_dtype == dtype, VRR_ERR_UNSUPPORTED, "gemm_ex (SIMT): the MUL epilogue needs c_dtype == dtype");
    if (int rc = gemm_simt(a, b, c, M, N, K, trans_a, trans_b, dtype, c_dtype, (cudaStream_t)stream)) return rc;
    return mul_inplace(c, c2, (size_t)M * N, c_dtype, (cudaStream_t)stream);
  }
  if (epilogue != VRR_EPI_NONE) {
    VRR_REQUIRE(c_dtype == dtype, VRR_ERR_UNSUPPORTED, "gemm_ex (SIMT): the bias epilogues need c_dtype == dtype");
    if (epilogue == VRR_EPI_BIAS_GELU_GRAD) {  // (h, gelu(h)) into (c2, c), then c2 <- gelu'(h)
      if (int rc = gemm_simt_bias(a, b, c2, c, bias, M, N, K, trans_a, trans_b, dtype, 1, (cudaStream_t)stream)) return rc;
      return gelu_grad_inplace(c2, (size_t)M * N, dtype, (cudaStream_t)stream);
    }
    return gemm_simt_bias(a, b, c, c2, bias, M, N, K, trans_a, trans_b, dtype,
                          epilogue == VRR_EPI_BIAS_GELU ? 1 : (epilogue == VRR_EPI_BIAS_GELU_ACT ? 2 : 0), (cudaStream_t)stream);
  }
  return gemm_simt(a, b, c, M, N, K, trans_a, trans_b, dtype, c_dtype, (cudaStream_t)stream);
}

int vrr_gemm_mul_colsum(const void* a, const void* b, void* c, const void* mul, float* col_sums, int M, int N, int K,
                        int trans_a, int trans_b, int dtype, void* stream) {
  VRR_REQUIRE(a && b && c && mul && col_sums, VRR_ERR_INVALID_ARG, "gemm_mul_colsum: NULL pointer");
  VRR_REQUIRE(M > 0 && N > 0 && K > 0 && N % 4 == 0, VRR_ERR_INVALID_ARG, "gemm_mul_colsum: bad sizes (N %% 4 == 0)");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "gemm_mul_colsum: bad dtype %d", dtype);
  if (int rc = require_device()) return rc;
  const int impl = g_impl.load();
  if (dtype == VRR_BF16 && impl != VRR_IMPL_SIMT && gemm_bf16_tc_supported(M, N, K, trans_a, trans_b, VRR_BF16, VRR_EPI_MUL)) {
    VRR_COUNT_FAMILY(VRR_IMPL_TCGEN05);
    VRR_CUDA(cudaMemsetAsync(col_sums, 0, (size_t)N * sizeof(float), (cudaStream_t)stream));
    return gemm_bf16_tc(a, b, c, const_cast<void*>(mul), nullptr, M, N, K, trans_a, trans_b, VRR_BF16, VRR_EPI_MUL, 0,
                        (cudaStream_t)stream, col_sums);
  }
  if (int rc = vrr_gemm_ex(a, b, c, const_cast<void*>(mul), nullptr, M, N, K, trans_a, trans_b, dtype, dtype, VRR_EPI_MUL, 0, stream))
    return rc;
  return colsum(c, col_sums, M, N, dtype, (cudaStream_t)stream);
}

int vrr_attn_fwd(const void* planes, const vrr_bias_desc* bias, void* out, float* lse, int B, int H, int N,
                 int Dh, float scale, int dtype, void* stream) {
  VRR_REQUIRE(planes && out && lse, VRR_ERR_INVALID_ARG, "attn_fwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "attn_fwd: bad dtype %d", dtype);
  VRR_REQUIRE(B > 0 && H > 0 && N > 0, VRR_ERR_INVALID_ARG, "attn_fwd: bad sizes");
  VRR_REQUIRE(dh_ok(Dh), VRR_ERR_UNSUPPORTED, "attn_fwd: head dim %d unsupported (16, 32, 64)", Dh);
  if (int rc = check_bias(bias, H, N)) return rc;
  if (int rc = require_device()) return rc;
  const int impl = g_impl.load();
  if (dtype == VRR_BF16 && impl != VRR_IMPL_SIMT && attn_fwd_tc_supported(B, H, N, Dh, bias)) {
    VRR_COUNT_FAMILY(VRR_IMPL_TCGEN05);
    const int variant = g_attn_fwd_variant.load();
    const bool no_bias = bias == nullptr || bias->mode == VRR_BIAS_NONE;
    if (variant >= 3 && !(variant >= 4 && no_bias) && attn_fwd_ws_supported(B, H, N, Dh, bias))  // short sequences with bias
      return attn_fwd_ws(planes, bias, out, lse, B, H, N, Dh, scale, (cudaStream_t)stream);
    return attn_fwd_tc(planes, bias, out, lse, B, H, N, Dh, scale, (cudaStream_t)stream);
  }
  VRR_REQUIRE(impl != VRR_IMPL_TCGEN05, VRR_ERR_UNSUPPORTED,
              "attn_fwd: tcgen05 kernel forced but shape/dtype unsupported (bf16, Dh=64)");
  VRR_COUNT_FAMILY(VRR_IMPL_SIMT);
  return attn_fwd_simt(planes, bias, out, lse, B, H, N, Dh, scale, dtype, (cudaStream_t)stream);
}

size_t vrr_attn_bwd_workspace_bytes(int B, int H, int N, int Dh, const vrr_bias_desc* bias) {
  (void)Dh;
  size_t delta = (size_t)B * H * N * sizeof(float);
  size_t lut = (size_t)H * (size_t)bias_lut_len(bias, N) * sizeof(float);
  return ((delta + 255) / 256) * 256 + ((lut + 255) / 256) * 256 + 256;
}

int vrr_attn_bwd(const void* planes, const vrr_bias_desc* bias, const void* out, const void* d_out,
                 const float* lse, void* d_planes, float* d_bias_param, void* workspace, size_t workspace_bytes,
                 int B, int H, int N, int Dh, float scale, int dtype, void* stream) {
  VRR_REQUIRE(planes && out && d_out && lse && d_planes && workspace, VRR_ERR_INVALID_ARG, "attn_bwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype), VRR_ERR_INVALID_ARG, "attn_bwd: bad dtype %d", dtype);
  VRR_REQUIRE(B > 0 && H > 0 && N > 0, VRR_ERR_INVALID_ARG, "attn_bwd: bad sizes");
  VRR_REQUIRE(dh_ok(Dh), VRR_ERR_UNSUPPORTED, "attn_bwd: head dim %d unsupported (16, 32, 64)", Dh);
  if (int rc = check_bias(bias, H, N)) return rc;
  const bool has_bias = bias && bias->mode != VRR_BIAS_NONE;
  VRR_REQUIRE(!has_bias || d_bias_param, VRR_ERR_INVALID_ARG, "attn_bwd: d_bias_param required with a bias");
  VRR_REQUIRE(workspace_bytes >= vrr_attn_bwd_workspace_bytes(B, H, N, Dh, bias), VRR_ERR_WORKSPACE,
              "attn_bwd: workspace %zu < %zu bytes", workspace_bytes, vrr_attn_bwd_workspace_bytes(B, H, N, Dh, bias));
  VRR_REQUIRE(((uintptr_t)workspace & 255) == 0, VRR_ERR_INVALID_ARG, "attn_bwd: workspace must be 256-byte aligned");
  if (int rc = require_device()) return rc;
  float* delta = (float*)workspace;
  size_t delta_bytes = (((size_t)B * H * N * sizeof(float) + 255) / 256) * 256;
  float* d_lut = (float*)((char*)workspace + delta_bytes);
  const int impl = g_impl.load();
  if (dtype == VRR_BF16 && impl != VRR_IMPL_SIMT && attn_bwd_tc2_supported(B, H, N, Dh, bias)) {
    VRR_COUNT_FAMILY(VRR_IMPL_TCGEN05);
    if (g_attn_bwd_variant.load() >= 3 && attn_bwd_ws_supported(B, H, N, Dh, bias))  // short sequences, no bias
      return attn_bwd_ws(planes, out, d_out, lse, d_planes, B, H, N, Dh, scale, (cudaStream_t)stream);
    return attn_bwd_tc2(planes, bias, out, d_out, lse, d_planes, d_bias_param, delta, B, H, N, Dh, scale,
                        (cudaStream_t)stream);
  }
  // e.g. polynomial degree > 3 (the tcgen05 kernel keeps 4 power sums per thread): explicit under
  // VRR_IMPL_TCGEN05, and visible to AUTO callers through vrr_family_count(VRR_IMPL_SIMT)
  VRR_REQUIRE(impl != VRR_IMPL_TCGEN05, VRR_ERR_UNSUPPORTED,
              "attn_bwd: tcgen05 kernel forced but shape/dtype unsupported (bf16, Dh=64, poly degree <= 3)");
  VRR_COUNT_FAMILY(VRR_IMPL_SIMT);
  return attn_bwd_simt(planes, bias, out, d_out, lse, d_planes, d_bias_param, delta, d_lut, B, H, N, Dh, scale,
                       dtype, (cudaStream_t)stream);
}

int vrr_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int M,
                      int E, float eps, int x_dtype, int y_dtype, void* stream) {
  VRR_REQUIRE(x && gamma && beta && y && mean && rstd, VRR_ERR_INVALID_ARG, "layernorm_fwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype), VRR_ERR_INVALID_ARG, "layernorm_fwd: bad dtype");
  VRR_REQUIRE(M > 0 && E > 0, VRR_ERR_INVALID_ARG, "layernorm_fwd: bad sizes");
  VRR_REQUIRE((((uintptr_t)x | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0, VRR_ERR_INVALID_ARG,
              "layernorm_fwd: pointers must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  return layernorm_fwd(x, gamma, beta, y, mean, rstd, M, E, eps, x_dtype, y_dtype, (cudaStream_t)stream);
}

int vrr_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                      void* dx, float* dgamma, float* dbeta, int M, int E, int x_dtype, int y_dtype, void* stream) {
  VRR_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, VRR_ERR_INVALID_ARG, "layernorm_bwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(x_dtype) && dtype_ok(y_dtype), VRR_ERR_INVALID_ARG, "layernorm_bwd: bad dtype");
  VRR_REQUIRE(M > 0 && E > 0, VRR_ERR_INVALID_ARG, "layernorm_bwd: bad sizes");
  VRR_REQUIRE((((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)gamma) & 15) == 0, VRR_ERR_INVALID_ARG,
              "layernorm_bwd: pointers must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  return layernorm_bwd(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, M, E, x_dtype, y_dtype, (cudaStream_t)stream);
}

int vrr_add_layernorm_fwd(const void* x, const void* branch, void* x_new, const float* gamma, const float* beta,
                          void* y, float* mean, float* rstd, int M, int E, float eps, int branch_dtype, int y_dtype,
                          void* stream) {
  VRR_REQUIRE(x && branch && x_new && gamma && beta && y && mean && rstd, VRR_ERR_INVALID_ARG, "add_layernorm_fwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(branch_dtype) && dtype_ok(y_dtype), VRR_ERR_INVALID_ARG, "add_layernorm_fwd: bad dtype");
  VRR_REQUIRE(M > 0 && E > 0, VRR_ERR_INVALID_ARG, "add_layernorm_fwd: bad sizes");
  VRR_REQUIRE((((uintptr_t)x | (uintptr_t)branch | (uintptr_t)x_new | (uintptr_t)y | (uintptr_t)gamma | (uintptr_t)beta) & 15) == 0,
              VRR_ERR_INVALID_ARG, "add_layernorm_fwd: pointers must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  return add_layernorm_fwd(x, branch, x_new, gamma, beta, y, mean, rstd, M, E, eps, branch_dtype, y_dtype,
                           (cudaStream_t)stream);
}

int vrr_add_layernorm_bwd(const void* dy, const void* d_xnew, const void* x_new, const float* gamma, const float* mean,
                          const float* rstd, void* dx, void* d_branch, float* dgamma, float* dbeta, int M, int E,
                          int branch_dtype, int y_dtype, void* stream) {
  VRR_REQUIRE(dy && x_new && gamma && mean && rstd && dx && d_branch && dgamma && dbeta, VRR_ERR_INVALID_ARG,
              "add_layernorm_bwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(branch_dtype) && dtype_ok(y_dtype), VRR_ERR_INVALID_ARG, "add_layernorm_bwd: bad dtype");
  VRR_REQUIRE(M > 0 && E > 0, VRR_ERR_INVALID_ARG, "add_layernorm_bwd: bad sizes");
  VRR_REQUIRE((((uintptr_t)dy | (uintptr_t)d_xnew | (uintptr_t)x_new | (uintptr_t)dx | (uintptr_t)d_branch | (uintptr_t)gamma) & 15) == 0,
              VRR_ERR_INVALID_ARG, "add_layernorm_bwd: pointers must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  return add_layernorm_bwd(dy, d_xnew, x_new, gamma, mean, rstd, dx, d_branch, dgamma, dbeta, M, E, branch_dtype,
                           y_dtype, (cudaStream_t)stream);
}

int vrr_colsum(const void* x, float* out, int M, int C, int dtype, void* stream) {
  VRR_REQUIRE(x && out, VRR_ERR_INVALID_ARG, "colsum: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype) && M > 0 && C > 0, VRR_ERR_INVALID_ARG, "colsum: bad arguments");
  VRR_REQUIRE(((uintptr_t)x & 15) == 0, VRR_ERR_INVALID_ARG, "colsum: x must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  return colsum(x, out, M, C, dtype, (cudaStream_t)stream);
}

int vrr_gelu_bwd(const void* dy, const void* h, void* dh, float* db, int M, int C, int dtype, void* stream) {
  VRR_REQUIRE(dy && h && dh && db, VRR_ERR_INVALID_ARG, "gelu_bwd: NULL pointer");
  VRR_REQUIRE(dtype_ok(dtype) && M > 0 && C > 0, VRR_ERR_INVALID_ARG, "gelu_bwd: bad arguments");
  VRR_REQUIRE((((uintptr_t)dy | (uintptr_t)h | (uintptr_t)dh) & 15) == 0, VRR_ERR_INVALID_ARG,
              "gelu_bwd: pointers must be 16-byte aligned");
  if (int rc = require_device()) return rc;
  return gelu_bwd_colsum(dy, h, dh, db, M, C, dtype, (cudaStream_t)stream);
}

}  // extern "C"
