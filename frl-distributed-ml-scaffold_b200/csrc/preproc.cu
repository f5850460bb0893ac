// K5 — device-side batch preprocessing and dtype/scale copies (sm_100a).
//
// The reference transforms every sample in Python on the host (reference transform.py:25-38,
// multitask_problem.py:56-71) and ships fp32 over a pageable copy.  Here the raw bytes are
// shipped once (u8 / f32 / bf16) and one streaming pass normalises and converts the whole
// batch: dst = src * scale[c] + bias[c], c = (i / inner) % channels.
#include "frl_common.cuh"

namespace frl {

constexpr int kPThreads = 256;

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<uint8_t>(uint8_t v) { return static_cast<float>(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename S, int N> struct alignas(sizeof(S) * N > 16 ? 16 : sizeof(S) * N) Pack { S v[N]; };

// 8 elements per thread: u8 -> 8 B load, bf16 -> 16 B, f32 -> 2 x 16 B; stores 16 B / 2 x 16 B.
template <typename S, typename D>
__global__ void __launch_bounds__(kPThreads)
affine_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n, int64_t inner,
              int64_t channels, const float* __restrict__ scale, const float* __restrict__ bias,
              float uniform_scale) {
    const int64_t n8 = n >> 3;
    const int64_t stride = static_cast<int64_t>(gridDim.x) * kPThreads;
    const bool per_channel = (scale != nullptr) || (bias != nullptr);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kPThreads + threadIdx.x; i < n8; i += stride) {
        const Pack<S, 8> in = *reinterpret_cast<const Pack<S, 8>*>(src + (i << 3));
        Pack<D, 8> out;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float x = to_f32<S>(in.v[k]);
            if (per_channel) {
                const int64_t c = (((i << 3) + k) / inner) % channels;
                x = fmaf(x, scale ? __ldg(scale + c) : 1.f, bias ? __ldg(bias + c) : 0.f);
            } else {
                x *= uniform_scale;
            }
            out.v[k] = from_f32<D>(x);
        }
        *reinterpret_cast<Pack<D, 8>*>(dst + (i << 3)) = out;
    }
    // tail
    if (blockIdx.x == 0) {
        const int64_t e = (n8 << 3) + threadIdx.x;
        if (threadIdx.x < 8 && e < n) {
            float x = to_f32<S>(src[e]);
            if (per_channel) {
                const int64_t c = (e / inner) % channels;
                x = fmaf(x, scale ? __ldg(scale + c) : 1.f, bias ? __ldg(bias + c) : 0.f);
            } else {
                x *= uniform_scale;
            }
            dst[e] = from_f32<D>(x);
        }
    }
}

template <typename S, typename D>
static int launch_affine(const void* src, void* dst, int64_t n, int64_t inner, int64_t channels,
                         const float* scale, const float* bias, float uscale, cudaStream_t st,
                         const char* name) {
    int64_t want = ((n >> 3) + kPThreads - 1) / kPThreads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    affine_kernel<S, D><<<static_cast<int>(want), kPThreads, 0, st>>>(
        static_cast<const S*>(src), static_cast<D*>(dst), n, inner, channels, scale, bias, uscale);
    return after_launch(name);
}

static int dispatch_affine(const void* src, int sd, void* dst, int dd, int64_t n, int64_t inner,
                           int64_t channels, const float* scale, const float* bias, float uscale,
                           cudaStream_t st, const char* name) {
    FRL_REQUIRE(n >= 0, FRL_E_ARG, "%s: n < 0", name);
    if (n == 0) return 0;
    FRL_REQUIRE(src && dst, FRL_E_ARG, "%s: null src/dst", name);
    FRL_REQUIRE(aligned16(src) && aligned16(dst), FRL_E_ALIGN, "%s: src/dst must be 16-byte aligned", name);
    FRL_REQUIRE(inner >= 1 && channels >= 1, FRL_E_ARG, "%s: inner/channels must be >= 1", name);
#define FRL_CASE(SD, DD, S, D) \
    if (sd == SD && dd == DD) return launch_affine<S, D>(src, dst, n, inner, channels, scale, bias, uscale, st, name)
    FRL_CASE(FRL_U8, FRL_F32, uint8_t, float);
    FRL_CASE(FRL_U8, FRL_BF16, uint8_t, __nv_bfloat16);
    FRL_CASE(FRL_F32, FRL_F32, float, float);
    FRL_CASE(FRL_F32, FRL_BF16, float, __nv_bfloat16);
    FRL_CASE(FRL_BF16, FRL_F32, __nv_bfloat16, float);
    FRL_CASE(FRL_BF16, FRL_BF16, __nv_bfloat16, __nv_bfloat16);
#undef FRL_CASE
    set_error("%s: unsupported dtype pair %d -> %d", name, sd, dd);
    return FRL_E_DTYPE;
}

}  // namespace frl

using namespace frl;

extern "C" int frl_preproc_affine(const void* src, int src_dtype, void* dst, int dst_dtype,
                                  int64_t n, int64_t inner, int64_t channels, const float* scale,
                                  const float* bias, void* stream) {
    return dispatch_affine(src, src_dtype, dst, dst_dtype, n, inner, channels, scale, bias, 1.f,
                           static_cast<cudaStream_t>(stream), "frl_preproc_affine");
}

extern "C" int frl_cast_scale(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n,
                              float scale, void* stream) {
    return dispatch_affine(src, src_dtype, dst, dst_dtype, n, 1, 1, nullptr, nullptr, scale,
                           static_cast<cudaStream_t>(stream), "frl_cast_scale");
}
