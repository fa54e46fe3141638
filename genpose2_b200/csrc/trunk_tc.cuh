// Tensor-core (tcgen05 + TMEM + TMA bulk copies) tile evaluator of the ScoreNet trunk, bf16 operands
// with fp32 accumulation.  One CTA = one 128-row tile = one UMMA M=128 accumulator:
//
//   warp 0      TMA producer : streams the 16 weight chunks (256 n x 64 k bf16, pre-swizzled images in
//                              global, 32 KB each, L2 resident) through a 3-stage shared-memory ring
//                              with cp.async.bulk + mbarrier complete_tx
//   warp 1      MMA issuer   : one lane issues tcgen05.mma (M=128, N=256, K=16) chains, accumulators
//                              in TMEM (2 x 256 fp32 columns), commits to mbarriers
//   warps 2..9  epilogue     : 256 threads = 128 rows x 2 column halves; layer 1 on CUDA cores
//                              (9 -> 256), TMEM -> registers (tcgen05.ld), bias / ReLU / bf16 pack
//                              into the swizzled A-operand buffer, and for the three heads
//                              + proj[obj] + tq, ReLU, 256 -> 3 output layer, row reduction
//
// Per evaluation: D1 = h1 . W2^T (4 chunks) ; h2 = relu(D1 + b2) ; D2_h = h2 . Whp_h^T (3 x 4 chunks).
// Operand layout: canonical K-major SWIZZLE_128B (8-row x 128-byte atoms, 16-byte chunk index XOR row%8).
#pragma once
#include <cuda_bf16.h>

#include "tc_ptx.cuh"
#include "trunk.cuh"

namespace gp {
namespace tc {

constexpr int RT = 128;               // rows per tile (UMMA M)
constexpr int NTHREADS = 320;         // 10 warps
constexpr int NSTAGE = 3;
constexpr int CHUNK_BYTES = 256 * 64 * 2;   // [256 n][64 k] bf16
constexpr int NCHUNK = 16;                  // 4 (W2) + 3 x 4 (heads)
constexpr int ATOM_BYTES = 128 * 128;       // A operand: [128 rows][64 k] bf16
constexpr uint32_t TMEM_COLS = 512;

struct Smem {
    uint8_t ring[NSTAGE][CHUNK_BYTES];  // must be 1024-byte aligned (struct is placed on a 1024 boundary)
    uint8_t abuf[4][ATOM_BYTES];        // h1 / h2 as A operand, 4 K-atoms
    float tq[6 * 768];
    float x[RT * 12];
    float out[2][RT * 12];
    float four[6 * 128];
    float tfeat[6 * 128];
    float times[8];
    double red[16];
    int obj[RT];
    unsigned long long full[NSTAGE], empty[NSTAGE], a_ready, d_full[2], d_free;
    uint32_t tmem_base;
};

// pipeline state carried across evaluations (every thread holds a copy, each role uses its own fields)
struct State {
    uint32_t loads = 0;      // producer: bulk copies issued so far
    uint32_t consumed = 0;   // MMA issuer: chunks consumed so far
    uint32_t a_phase = 0;    // MMA issuer: parity of the next a_ready completion
    uint32_t dfree_phase = 0;
    uint32_t dfull_phase[2] = {0, 0};  // epilogue: parity of the next d_full[i] completion
};

// store 8 consecutive bf16 (columns n0..n0+7 of `row`) into the swizzled A buffer
__device__ __forceinline__ void store_a8(Smem &S, int row, int n0, const float *v) {
    const int atom = n0 >> 6, c16 = (n0 & 63) >> 3;
    uint4 pk;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
    pk.x = *reinterpret_cast<uint32_t *>(&p0); pk.y = *reinterpret_cast<uint32_t *>(&p1);
    pk.z = *reinterpret_cast<uint32_t *>(&p2); pk.w = *reinterpret_cast<uint32_t *>(&p3);
    *reinterpret_cast<uint4 *>(&S.abuf[atom][row * 128 + ((c16 ^ (row & 7)) << 4)]) = pk;
}

// one-time setup / teardown (all threads call)
__device__ __forceinline__ void setup(Smem &S, State &st, const float *__restrict__ P) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        mbar_init(&S.a_ready, 8);
        mbar_init(&S.d_full[0], 1);
        mbar_init(&S.d_full[1], 1);
        mbar_init(&S.d_free, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) tmem_alloc(&S.tmem_base, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0 && (tid & 31) == 0) {  // prefill the ring
        const uint8_t *src = reinterpret_cast<const uint8_t *>(P + TrunkLayout::W_TC);
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_arrive_expect_tx(&S.full[s], CHUNK_BYTES);
            bulk_g2s(S.ring[s], src + (size_t)s * CHUNK_BYTES, CHUNK_BYTES, &S.full[s]);
        }
        st.loads = NSTAGE;
    }
}

__device__ __forceinline__ void teardown(Smem &S, State &st) {
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 1 && (tid & 31) == 0) {  // drain the NSTAGE loads that are still in flight
        for (int i = 0; i < NSTAGE; ++i) {
            const uint32_t g = st.consumed + i;
            mbar_wait(&S.full[g % NSTAGE], (g / NSTAGE) & 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(S.tmem_base, TMEM_COLS);
}

// f_theta for the 128 rows in S.x -> S.out[0][r*12 + c]; ends with __syncthreads().
__device__ __forceinline__ void forward(const float *__restrict__ P, const float *__restrict__ proj, Smem &S,
                                        State &st, const float *s_tq) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = S.tmem_base;
    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            const uint8_t *src = reinterpret_cast<const uint8_t *>(P + TrunkLayout::W_TC);
            for (int i = 0; i < NCHUNK; ++i) {
                const uint32_t L = st.loads;
                const uint32_t s = L % NSTAGE;
                mbar_wait(&S.empty[s], ((L / NSTAGE) + 1) & 1);
                mbar_arrive_expect_tx(&S.full[s], CHUNK_BYTES);
                bulk_g2s(S.ring[s], src + (size_t)(L % NCHUNK) * CHUNK_BYTES, CHUNK_BYTES, &S.full[s]);
                st.loads = L + 1;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t a_base = smem_u32(&S.abuf[0][0]);
            auto gemm = [&](uint32_t dcol) {
                for (int kc = 0; kc < 4; ++kc) {
                    const uint32_t g = st.consumed;
                    const uint32_t s = g % NSTAGE;
                    mbar_wait(&S.full[s], (g / NSTAGE) & 1);
                    tc_fence_after();
                    const uint32_t b_base = smem_u32(&S.ring[s][0]);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma_bf16(tmem + dcol, make_desc(a_base + kc * ATOM_BYTES + kk * 32), make_desc(b_base + kk * 32),
                                  kIdesc, (kc | kk) ? 1u : 0u);
                    umma_commit(&S.empty[s]);
                    st.consumed = g + 1;
                }
            };
            mbar_wait(&S.a_ready, st.a_phase); st.a_phase ^= 1;   // h1 ready
            tc_fence_after();
            gemm(0);
            umma_commit(&S.d_full[0]);
            mbar_wait(&S.a_ready, st.a_phase); st.a_phase ^= 1;   // h2 ready (and D1 drained)
            tc_fence_after();
            gemm(256);                                            // head 0 -> cols 256..511
            umma_commit(&S.d_full[1]);
            gemm(0);                                              // head 1 -> cols 0..255
            umma_commit(&S.d_full[0]);
            mbar_wait(&S.d_free, st.dfree_phase); st.dfree_phase ^= 1;  // head 0 drained from cols 256..511
            tc_fence_after();
            gemm(256);                                            // head 2
            umma_commit(&S.d_full[1]);
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps ----------------
        const int e = warp - 2;
        const int row = 32 * (warp & 3) + lane;   // TMEM lane quarter is fixed by warp % 4
        const int half = e >> 2;                  // column half 0 / 1
        const int c0 = half * 128;
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        // layer 1 (9 -> 256) for this row, columns c0..c0+127
        {
            float xv[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) xv[k] = S.x[row * 12 + k];
            for (int n0 = c0; n0 < c0 + 128; n0 += 8) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(P + TrunkLayout::B1 + n0 + j);
#pragma unroll
                for (int k = 0; k < 9; ++k) {
                    const float4 wa = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::W1T + k * 256 + n0));
                    const float4 wb = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::W1T + k * 256 + n0 + 4));
                    v[0] = fmaf(xv[k], wa.x, v[0]); v[1] = fmaf(xv[k], wa.y, v[1]);
                    v[2] = fmaf(xv[k], wa.z, v[2]); v[3] = fmaf(xv[k], wa.w, v[3]);
                    v[4] = fmaf(xv[k], wb.x, v[4]); v[5] = fmaf(xv[k], wb.y, v[5]);
                    v[6] = fmaf(xv[k], wb.z, v[6]); v[7] = fmaf(xv[k], wb.w, v[7]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                store_a8(S, row, n0, v);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.a_ready);
        }
        // layer 2 epilogue: h2 = relu(D1 + b2) -> A buffer
        {
            mbar_wait(&S.d_full[0], st.dfull_phase[0]); st.dfull_phase[0] ^= 1;
            tc_fence_after();
            for (int g = 0; g < 4; ++g) {
                uint32_t r[32];
                tmem_ld32(lane_addr + c0 + g * 32, r);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        v[j] = fmaxf(__uint_as_float(r[j8 * 8 + j]) + __ldg(P + TrunkLayout::B2 + c0 + g * 32 + j8 * 8 + j), 0.f);
                    store_a8(S, row, c0 + g * 32 + j8 * 8, v);
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.a_ready);
        }
        // heads
        float acc[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) acc[c] = 0.f;
        const int o = S.obj[row];
        const float *prow = proj + (size_t)(o < 0 ? 0 : o) * 768;
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            const int buf = (h == 1) ? 0 : 1;
            mbar_wait(&S.d_full[buf], st.dfull_phase[buf]); st.dfull_phase[buf] ^= 1;
            tc_fence_after();
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            for (int g = 0; g < 4; ++g) {
                uint32_t r[32];
                tmem_ld32(lane_addr + (buf ? 256 : 0) + c0 + g * 32, r);
                const int nb = h * 256 + c0 + g * 32;
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 pj = __ldg(reinterpret_cast<const float4 *>(prow + nb + j4 * 4));
                    const float4 tq = *reinterpret_cast<const float4 *>(s_tq + nb + j4 * 4);
                    const float z0 = fmaxf(__uint_as_float(r[j4 * 4 + 0]) + (pj.x + tq.x), 0.f);
                    const float z1 = fmaxf(__uint_as_float(r[j4 * 4 + 1]) + (pj.y + tq.y), 0.f);
                    const float z2 = fmaxf(__uint_as_float(r[j4 * 4 + 2]) + (pj.z + tq.z), 0.f);
                    const float z3 = fmaxf(__uint_as_float(r[j4 * 4 + 3]) + (pj.w + tq.w), 0.f);
                    const float4 w0 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(nb + j4 * 4 + 0) * 4));
                    const float4 w1 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(nb + j4 * 4 + 1) * 4));
                    const float4 w2 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(nb + j4 * 4 + 2) * 4));
                    const float4 w3 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(nb + j4 * 4 + 3) * 4));
                    a0 = fmaf(z0, w0.x, a0); a1 = fmaf(z0, w0.y, a1); a2 = fmaf(z0, w0.z, a2);
                    a0 = fmaf(z1, w1.x, a0); a1 = fmaf(z1, w1.y, a1); a2 = fmaf(z1, w1.z, a2);
                    a0 = fmaf(z2, w2.x, a0); a1 = fmaf(z2, w2.y, a1); a2 = fmaf(z2, w2.z, a2);
                    a0 = fmaf(z3, w3.x, a0); a1 = fmaf(z3, w3.y, a1); a2 = fmaf(z3, w3.z, a2);
                }
            }
            acc[h * 3 + 0] = a0; acc[h * 3 + 1] = a1; acc[h * 3 + 2] = a2;
            if (h == 0) {  // cols 256..511 may now be overwritten by head 2
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.d_free);
            }
        }
        tc_fence_before();
#pragma unroll
        for (int c = 0; c < 9; ++c) S.out[half][row * 12 + c] = acc[c];
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
        if (half == 0) {
#pragma unroll
            for (int c = 0; c < 9; ++c)
                S.out[0][row * 12 + c] = (S.out[0][row * 12 + c] + S.out[1][row * 12 + c]) + __ldg(P + TrunkLayout::BO + c);
        }
    }
    __syncthreads();
}

}  // namespace tc
}  // namespace gp
