// out[M,N] = A[M,K] * Bt[N,K]^T (+ bias) on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32, accumulators in
// TMEM, weights by TMA.  This is the consumer of the Wigner action: the first layer of the reference's
// DeconvNet, ConvTranspose2d(M*C -> hidden, 4, 1, 0) on a 1x1 input (experiments/nets.py:65-66), is the GEMM
// (N_samples, 810) x (810, 16*hidden), and ActionNet's optional MLP (decoders.py:39-41,58-59) starts with a Linear of the same
// shape -- the only dense contractions on the hot path.
//
// Why the action is NOT produced inside this kernel: one 128-sample A tile is 128 x 810 x 4 B = 415 KB
// (shared memory holds 227 KB) and the accumulators of all 16*hidden = 3 200 output columns are 1.6 MB (TMEM holds 256 KB =
// 512 columns), so an in-kernel producer would have to recompute the Wigner chain once per 512-column slice of the output
// (7x the FP32 work of a kernel that already runs at the roofline), and the weight gradient needs y in memory anyway.  What
// removes the HBM round trip of y is granularity: the host interleaves the Wigner forward and this GEMM in chunks of samples
// whose y (3 240 B/sample) stays in the 126 MB L2 (lv_action_gemm_fwd_f32), so A is read from L2, not from HBM.
//
// Precision: TF32 operands (10-bit mantissa), FP32 accumulation -- the arithmetic cuDNN uses for the reference's FP32
// ConvTranspose2d on this GPU under PyTorch's defaults (torch.backends.cudnn.allow_tf32 = True).  Held to the float64
// oracle at 2e-3 relative to the output's rms (tests/test_gpu_gemm.py); the FP32-exact path stays the default.
//
// Anatomy (one 128 x 256 output tile per CTA -- 24 operand bytes per MFLOP from L2 instead of the 32 of a square 128 tile, which
// was the limit --, K in blocks of 32 floats = one 128-byte swizzle row, 4-stage ring of 16 KB A + 32 KB weight tiles):
//   warps 0-3  A producers: 8-byte cp.async pieces (a warp covers two whole row segments) placed in the SWIZZLE_128B K-major layout by hand (rows
//              of y are 3 240 B apart -- 8-byte, not 16-byte aligned, so TMA cannot address them), zero-filled past M and K;
//              thread 0 also issues the TMA load of the 256 x 32 weight tile (CU_TENSOR_MAP_SWIZZLE_128B, OOB rows/cols = 0).
//              After the main loop the same warps are the epilogue: tcgen05.ld 32 lanes x 32 columns, + bias, 128-bit stores.
//   warp 4     MMA issuer: waits full[s], one lane issues 4 x tcgen05.mma (M 128, N 256, K 8) per stage, tcgen05.commit -> empty[s];
//              after the last block commit -> tmem_full.  Owns the TMEM allocation (256 columns).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace lv {

constexpr int GT_BM = 128, GT_BK = 32, GT_STAGES = 4, GT_LOOKAHEAD = 2, GT_THREADS = 160;
constexpr uint32_t GT_TILE_BYTES = GT_BM * GT_BK * 4;          // A stage, 16 KB: 128 rows x 128 B
// BN = 256: 265 TFLOP/s at 65 536 x 3 200 x 810 (BN = 128: 168; BN = 256 with three stages: 205; cuBLAS TF32: 133); BN = 128 is kept
// for problems too small to fill the SMs with 256-wide tiles (lv_gemm_tf32_f32 chooses)
template <int BN> struct GtGeo {
    static constexpr uint32_t BTILE_BYTES = BN * GT_BK * 4;     // weight stage: BN rows x 128 B
    static constexpr size_t SMEM = size_t(GT_STAGES) * (GT_TILE_BYTES + BTILE_BYTES) + 1024 /* alignment slack */ + 256 /* barriers */;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
    // N >> 3 in [17,23), M >> 4 in [24,29)
    static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(BN >> 3) << 17) | (uint32_t(GT_BM >> 4) << 24);
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte offset = 8 rows x 128 B = 1024 >> 4 in [32,46),
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}
template <int GT_BN>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tm_b, const float* __restrict__ A, int64_t lda, const float* __restrict__ bias,
                 int bias_div, float* __restrict__ out, int64_t ldo, int M, int N, int K) {
    constexpr uint32_t GT_BTILE_BYTES = GtGeo<GT_BN>::BTILE_BYTES, GT_IDESC = GtGeo<GT_BN>::IDESC;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // swizzle atoms are 1024-byte aligned
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    uint8_t* sA = smem;                                                     // [STAGES][128 rows][128 B]
    uint8_t* sB = smem + GT_STAGES * GT_TILE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GT_STAGES * (GT_TILE_BYTES + GT_BTILE_BYTES));
    uint64_t* empty = full + GT_STAGES;
    uint64_t* tmem_full = empty + GT_STAGES;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(tmem_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = (N + GT_BN - 1) / GT_BN;                            // n fastest: the CTAs of one A tile run together
    const int m0 = int(blockIdx.x / n_tiles) * GT_BM, n0 = int(blockIdx.x % n_tiles) * GT_BN;
    const int nkb = (K + GT_BK - 1) / GT_BK;
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; ++s) { mbar_init(full + s, 129); mbar_init(empty + s, 1); }    // 128 A rows + the TMA's expect_tx
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_b)) : "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(uint32_t(GT_BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;

    if (warp < 4) {
        // ------------------------------------------------------------ A producers (+ thread 0: TMA of the weight tile)
        // Piece (i, tid): 8 bytes of row i*8 + tid/16 at byte offset (tid%16)*8 -- a warp instruction covers two whole 128-byte row
        // segments (coalesced), and the same thread later rounds exactly the pieces it copied (no hand-over between threads).
        const int prow = tid >> 4, pcol = tid & 15;
        const uint32_t poff = uint32_t(pcol & 1) * 8u;
        for (int kb = 0; kb < nkb + GT_LOOKAHEAD; ++kb) {
            if (kb < nkb) {
                const int s = kb % GT_STAGES;
                mbar_wait(empty + s, (uint32_t(kb / GT_STAGES) & 1u) ^ 1u);
                if (tid == 0) {
                    mbar_expect_tx(full + s, GT_BTILE_BYTES);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 :: "r"(smem_u32(sB + s * GT_BTILE_BYTES)), "l"(reinterpret_cast<uint64_t>(&tm_b)), "r"(kb * GT_BK), "r"(n0),
                                    "r"(smem_u32(full + s)) : "memory");
                }
                const uint32_t dtile = smem_u32(sA + s * GT_TILE_BYTES);
                const int kk = kb * GT_BK + 2 * pcol;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = i * 8 + prow, gr = m0 + r;
                    int nbytes = gr < M ? (K - kk) * 4 : 0;
                    nbytes = nbytes < 0 ? 0 : (nbytes > 8 ? 8 : nbytes);
                    // 16-byte chunk c = pcol / 2 of row r lands at chunk c ^ (r & 7) (SWIZZLE_128B)
                    const uint32_t dst = dtile + uint32_t(r) * 128u + ((uint32_t(pcol >> 1) ^ uint32_t(r & 7)) << 4) + poff;
                    const float* src = A + (nbytes > 0 ? int64_t(gr) * lda + kk : 0);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            if (kb >= GT_LOOKAHEAD) {
                asm volatile("cp.async.wait_group %0;" ::"n"(GT_LOOKAHEAD) : "memory");      // this thread's pieces of block kb - LOOKAHEAD have landed
                // The tensor core reads FP32 words and ignores the low 13 mantissa bits (truncation: every product shrinks by
                // 2^-11 on average).  Round the operands to TF32 (nearest, ties away: cvt.rna) in place, as cuBLAS / cuDNN do
                // before their TF32 MMAs.  (The weight tile arrives by TMA: the caller rounds Bt once, lv_round_tf32_f32.)
                const uint32_t rtile = smem_u32(sA + ((kb - GT_LOOKAHEAD) % GT_STAGES) * GT_TILE_BYTES);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int r = i * 8 + prow;
                    const uint32_t addr = rtile + uint32_t(r) * 128u + ((uint32_t(pcol >> 1) ^ uint32_t(r & 7)) << 4) + poff;
                    uint32_t x0, x1;
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(x0), "=r"(x1) : "r"(addr));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(x0) : "f"(__uint_as_float(x0)));
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(x1) : "f"(__uint_as_float(x1)));
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x0), "r"(x1) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                  // generic-proxy writes -> visible to the tensor core
                mbar_arrive(full + (kb - GT_LOOKAHEAD) % GT_STAGES);
            }
        }
        const int row = tid, grow = m0 + row;                // epilogue: thread t owns accumulator row (TMEM lane) t
        const bool valid = grow < M;
        // ------------------------------------------------------------ epilogue: TMEM -> registers -> + bias -> global
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float* orow = out + int64_t(grow) * ldo + n0;
        const bool vec = ((reinterpret_cast<uintptr_t>(out) | uintptr_t(ldo * 4)) & 15u) == 0;
#pragma unroll 1
        for (int c4 = 0; c4 < GT_BN / 32; ++c4) {
            uint32_t r[32];
            const uint32_t taddr = tmem + (uint32_t(warp * 32) << 16) + uint32_t(c4 * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                           "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                           "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                           "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (valid) {
                const int nb = n0 + c4 * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int n = nb + j + q;
                        v[q] = __uint_as_float(r[j + q]) + ((bias != nullptr && n < N) ? __ldg(bias + n / bias_div) : 0.f);
                    }
                    if (vec && nb + j + 3 < N) {
                        *reinterpret_cast<float4*>(orow + c4 * 32 + j) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (nb + j + q < N) orow[c4 * 32 + j + q] = v[q];
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ------------------------------------------------------------ MMA issuer
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % GT_STAGES;
            mbar_wait(full + s, uint32_t(kb / GT_STAGES) & 1u);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t da = umma_desc_sw128(smem_u32(sA + s * GT_TILE_BYTES));
                const uint64_t db = umma_desc_sw128(smem_u32(sB + s * GT_BTILE_BYTES));
#pragma unroll
                for (int k = 0; k < GT_BK / 8; ++k) {              // UMMA_K = 8 tf32 = 32 bytes: +2 in the (address >> 4) field
                    const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                                 :: "r"(tmem), "l"(da + uint64_t(2 * k)), "l"(db + uint64_t(2 * k)), "r"(GT_IDESC), "r"(acc) : "memory");
                }
                // frees the stage once the MMAs above have read it (tcgen05.commit implies fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(empty + s)) : "memory");
                if (kb == nkb - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(tmem_full)) : "memory");
            }
            __syncwarp();
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(uint32_t(GT_BN)) : "memory");
    }
}

// ---- host: tensor map of the K-major weight matrix Bt (N rows x K floats, row stride ldb floats) -------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace lv

namespace lv {
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
    const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (i < n) {
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(in[i]));
        out[i] = __uint_as_float(r);
    }
}
}  // namespace lv

// ====================================================================== C ABI
extern "C" int lv_round_tf32_f32(const float* in, float* out, int64_t n, void* stream) {
    if (n < 0) { lv::set_error("round_tf32: negative size"); return LV_ERR_ARG; }
    if (n == 0) return LV_OK;
    if (!in || !out) { lv::set_error("round_tf32: null pointer"); return LV_ERR_ARG; }
    lv::round_tf32_kernel<<<unsigned((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, n);
    return lv::check_launch("round_tf32");
}

extern "C" int lv_gemm_tf32_f32(const float* A, int64_t lda, const float* Bt, int64_t ldb, const float* bias, int bias_div, float* out,
                                int64_t ldo, int64_t M, int N, int K, void* stream) {
    if (M < 0 || N <= 0 || K <= 0 || lda < K || ldb < K || ldo < N || bias_div < 1 || M > 0x7fffffffLL) { lv::set_error("gemm_tf32: bad sizes"); return LV_ERR_ARG; }
    if (M == 0) return LV_OK;
    if (!A || !Bt || !out) { lv::set_error("gemm_tf32: null pointer"); return LV_ERR_ARG; }
    if ((reinterpret_cast<uintptr_t>(A) & 7u) || (lda & 1)) { lv::set_error("gemm_tf32: A must be 8-byte aligned with an even row stride"); return LV_ERR_ALIGN; }
    if ((reinterpret_cast<uintptr_t>(Bt) & 15u) || (ldb & 3)) { lv::set_error("gemm_tf32: Bt must be 16-byte aligned with a row stride that is a multiple of 4 floats"); return LV_ERR_ALIGN; }
    lv::EncodeTiledFn enc = lv::encode_tiled_fn();
    if (!enc) { lv::set_error("gemm_tf32: cuTensorMapEncodeTiled is not available from this driver"); return LV_ERR_UNSUPPORTED; }
    // tile width: 256 unless the output is at most 128 columns wide or 256-wide tiles would leave most SMs without a tile (small
    // dgrad-shaped problems)
    const int64_t mt = (M + lv::GT_BM - 1) / lv::GT_BM;
    int bn = 256;
    if (N <= 128 || mt * ((N + 255) / 256) < 64) bn = 128;
    if (const char* e = getenv("LV_GEMM_BN")) { if (atoi(e) == 128 || atoi(e) == 256) bn = atoi(e); }     // A/B runs
    CUtensorMap tm;
    const cuuint64_t gdim[2] = {cuuint64_t(K), cuuint64_t(N)};
    const cuuint64_t gstride[1] = {cuuint64_t(ldb) * 4};
    const cuuint32_t box[2] = {cuuint32_t(lv::GT_BK), cuuint32_t(bn)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(Bt), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { lv::set_error("gemm_tf32: cuTensorMapEncodeTiled failed (%d)", int(r)); return LV_ERR_ARG; }
    static bool opted[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !opted[dev]) {
        cudaError_t e = cudaFuncSetAttribute(lv::gemm_tf32_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lv::GtGeo<256>::SMEM));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(lv::gemm_tf32_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(lv::GtGeo<128>::SMEM));
        if (e != cudaSuccess) { lv::set_error("gemm_tf32: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return int(e); }
        opted[dev] = true;
    }
    const int64_t tiles = ((N + bn - 1) / bn) * mt;
    if (tiles > 0x7fffffffLL) { lv::set_error("gemm_tf32: too many tiles"); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (bn == 256)
        lv::gemm_tf32_kernel<256><<<unsigned(tiles), lv::GT_THREADS, lv::GtGeo<256>::SMEM, st>>>(tm, A, lda, bias, bias_div, out, ldo, int(M), N, K);
    else
        lv::gemm_tf32_kernel<128><<<unsigned(tiles), lv::GT_THREADS, lv::GtGeo<128>::SMEM, st>>>(tm, A, lda, bias, bias_div, out, ldo, int(M), N, K);
    return lv::check_launch("gemm_tf32");
}
