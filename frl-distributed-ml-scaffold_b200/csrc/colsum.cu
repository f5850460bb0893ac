// K6 — column sum of a row-major matrix: out[c] (+)= sum_r x[r, c]   (sm_100a).
//
// This is the bias gradient of a linear layer (db = sum over the batch of dY).  Stock autograd
// computes it with a generic reduction that re-reads dY at a fraction of HBM speed
// (at::reduce_kernel: ~27 us for a 4096x4096 bf16 dY, 5 launches per step in the MLP config);
// here it is one bandwidth-bound pass whose result lands directly in the gradient arena.
//
// Grid (column tiles, row splits).  A warp reads one row segment of 32 lanes x 8 columns with a
// single 16-byte (bf16) / two 16-byte (fp32) loads per lane, rows strided across the warps and
// row splits; per-CTA partials go to scratch and the last CTA of each column tile (atomic
// ticket) folds the splits in a fixed order — deterministic, no float atomics.
#include "frl_common.cuh"

namespace frl {

constexpr int kSThreads = 256;
constexpr int kSWarps = kSThreads / 32;
constexpr int kSCols = 32 * 8;          // columns per CTA tile
constexpr int kSMaxSplits = 64;

struct ColsumScratchHeader { unsigned int ticket[1]; };

template <typename T> __device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const f32x4 a = ld_stream_ro(reinterpret_cast<const f32x4*>(p));
    const f32x4 b = ld_stream_ro(reinterpret_cast<const f32x4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const bf16x8 r = ld_stream_ro(reinterpret_cast<const bf16x8*>(p));
    v[0] = bf16lo(r.a); v[1] = bf16hi(r.a); v[2] = bf16lo(r.b); v[3] = bf16hi(r.b);
    v[4] = bf16lo(r.c); v[5] = bf16hi(r.c); v[6] = bf16lo(r.d); v[7] = bf16hi(r.d);
}
template <typename T> __device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// partial layout: [tile][split][kSCols] floats, after the per-tile tickets
template <typename XT, typename OT, bool VEC>
__global__ void __launch_bounds__(kSThreads)
colsum_kernel(const XT* __restrict__ x, int64_t rows, int64_t cols, OT* __restrict__ out,
              int accumulate, unsigned int* __restrict__ tickets, float* __restrict__ partial) {
    __shared__ float sm[kSWarps][kSCols];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x, split = blockIdx.y, nsplit = gridDim.y;
    const int64_t c0 = static_cast<int64_t>(tile) * kSCols + lane * 8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (VEC) {
        if (c0 < cols) {      // cols % 8 == 0 on this path, so the 8 columns are all valid
            for (int64_t r = static_cast<int64_t>(split) * kSWarps + warp; r < rows;
                 r += static_cast<int64_t>(nsplit) * kSWarps) {
                float v[8];
                load8<XT>(x + r * cols + c0, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += v[k];
            }
        }
    } else {
        for (int64_t r = static_cast<int64_t>(split) * kSWarps + warp; r < rows;
             r += static_cast<int64_t>(nsplit) * kSWarps) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c0 + k < cols) acc[k] += ld1<XT>(x + r * cols + c0 + k);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) sm[warp][lane * 8 + k] = acc[k];
    __syncthreads();
    // fold the warps of this CTA: thread t owns column t of the tile
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kSWarps; ++w) s += sm[w][threadIdx.x];
    float* my = partial + (static_cast<int64_t>(tile) * nsplit + split) * kSCols;
    my[threadIdx.x] = s;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&tickets[tile], 1u) == static_cast<unsigned int>(nsplit - 1));
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int64_t c = static_cast<int64_t>(tile) * kSCols + threadIdx.x;
    if (c < cols) {
        float tot = 0.f;
        const float* base = partial + static_cast<int64_t>(tile) * nsplit * kSCols + threadIdx.x;
        for (int sp = 0; sp < nsplit; ++sp) tot += __ldcg(base + static_cast<int64_t>(sp) * kSCols);
        if (accumulate) tot += ld1<OT>(out + c);
        st1<OT>(out + c, tot);
    }
    if (threadIdx.x == 0) tickets[tile] = 0;
}

static inline int64_t colsum_tiles(int64_t cols) { return (cols + kSCols - 1) / kSCols; }
static inline int colsum_splits(int64_t rows, int64_t tiles) {
    int64_t want = (static_cast<int64_t>(sm_count()) * 4 + tiles - 1) / tiles;    // ~4 CTAs per SM
    const int64_t max_by_rows = (rows + kSWarps - 1) / kSWarps;
    if (want > max_by_rows) want = max_by_rows;
    if (want > kSMaxSplits) want = kSMaxSplits;
    if (want < 1) want = 1;
    return static_cast<int>(want);
}

}  // namespace frl

using namespace frl;

extern "C" int64_t frl_colsum_scratch_bytes(int64_t rows, int64_t cols) {
    if (rows < 0 || cols < 1) return -1;
    const int64_t tiles = colsum_tiles(cols);
    return tiles * static_cast<int64_t>(sizeof(unsigned int)) + 16 +
           tiles * kSMaxSplits * kSCols * static_cast<int64_t>(sizeof(float));
}

extern "C" int frl_colsum(const void* x, int x_dtype, int64_t rows, int64_t cols, void* out,
                          int out_dtype, int accumulate, void* scratch, void* stream) {
    FRL_REQUIRE(x && out && scratch && rows >= 0 && cols >= 1, FRL_E_ARG, "frl_colsum: bad args");
    FRL_REQUIRE((x_dtype == FRL_F32 || x_dtype == FRL_BF16) && (out_dtype == FRL_F32 || out_dtype == FRL_BF16),
                FRL_E_DTYPE, "frl_colsum: dtype");
    const int64_t tiles = colsum_tiles(cols);
    const int splits = colsum_splits(rows, tiles);
    unsigned int* tickets = static_cast<unsigned int*>(scratch);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) +
                                              ((tiles * sizeof(unsigned int) + 15) / 16) * 16);
    const bool vec = (cols % 8 == 0) && aligned16(x);
    dim3 grid(static_cast<unsigned int>(tiles), static_cast<unsigned int>(splits));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FRL_CS(XT, OT)                                                                              \
    do {                                                                                            \
        if (vec) colsum_kernel<XT, OT, true><<<grid, kSThreads, 0, st>>>(                          \
                static_cast<const XT*>(x), rows, cols, static_cast<OT*>(out), accumulate, tickets, partial); \
        else colsum_kernel<XT, OT, false><<<grid, kSThreads, 0, st>>>(                             \
                static_cast<const XT*>(x), rows, cols, static_cast<OT*>(out), accumulate, tickets, partial); \
    } while (0)
    if (x_dtype == FRL_F32 && out_dtype == FRL_F32) FRL_CS(float, float);
    else if (x_dtype == FRL_BF16 && out_dtype == FRL_BF16) FRL_CS(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == FRL_BF16 && out_dtype == FRL_F32) FRL_CS(__nv_bfloat16, float);
    else FRL_CS(float, __nv_bfloat16);
#undef FRL_CS
    return after_launch("frl_colsum");
}
