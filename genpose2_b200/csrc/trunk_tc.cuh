// Tensor-core (tcgen05 + TMEM + TMA bulk copies) tile evaluator of the ScoreNet trunk.
// One CTA = one 128-row tile = one UMMA M=128 accumulator, 10 warps:
//
//   warp 0      TMA producer : streams the 32 weight chunks of an evaluation ([128 n][64 k] bf16 images,
//                              pre-swizzled in global, 16 KB each, L2 resident) through a shared-memory
//                              ring with cp.async.bulk + mbarrier complete_tx, running ahead across
//                              layers and evaluations
//   warp 1      MMA issuer   : one lane issues tcgen05.mma (M128 N128 K16) chains, accumulators in
//                              TMEM (2 x 256 fp32 columns), commits to mbarriers
//   warps 2..9  epilogue     : 256 threads.  Layer 1 (9 -> 256) on CUDA cores, one output column per
//                              thread with its weights in registers; TMEM -> registers (tcgen05.ld), bias,
//                              ReLU, bf16 (hi / lo) re-pack into the swizzled A-operand buffers; heads:
//                              + (proj + tq) from shared memory, ReLU, 256 -> 3 output layer.
//
// NPASS = 1 ("bf16" mode): operands rounded to bf16.
// NPASS = 3 ("fp32" mode): every operand is split x = hi + lo (two bf16) and each product is accumulated as
//   hi*hi + lo*hi + hi*lo in fp32 (TMEM): 16 mantissa bits per operand.  On the reference's own fixtures this
//   is indistinguishable from a plain fp32 evaluation (3e-6 rad / 8e-7 at T0 = 0.55, the same as changing the
//   fp32 summation order), while plain bf16 costs 1e-3 rad / 3e-4.
//
// Per evaluation: D1 = h1 . W2^T ; h2 = relu(D1 + b2) ; D2_h = h2 . Whp_h^T (3 heads).  Operand layout:
// canonical K-major SWIZZLE_128B (8-row x 128-byte atoms, 16-byte chunk index XOR row % 8).
#pragma once
#include <cuda_bf16.h>

#include "tc_ptx.cuh"
#include "trunk.cuh"

namespace gp {
namespace tc {

constexpr int RT = 128;               // rows per tile (UMMA M)
constexpr int NTHREADS = 320;         // 10 warps
constexpr int IMG_BYTES = 128 * 128;  // one weight image: [128 n][64 k] bf16
constexpr int NCHUNK = 32;            // 4 GEMMs x 4 k-atoms x 2 n-halves
constexpr int ATOM_BYTES = 128 * 128; // A operand atom: [128 rows][64 k] bf16
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_SLOTS = 5;          // objects a tile may span for the shared-memory (proj + tq) table
constexpr uint32_t kIdescN128 = make_idesc_bf16(128, 128);

template <int NPASS>
struct Smem {
    static constexpr int IMAGES = NPASS == 3 ? 2 : 1;
    static constexpr int NSTAGE = NPASS == 3 ? 2 : 4;
    uint8_t ring[NSTAGE][IMAGES][IMG_BYTES];  // 64 KB; the struct sits on a 1024-byte boundary
    uint8_t abuf[IMAGES][4][ATOM_BYTES];      // h1 / h2 as A operand (hi, lo), 4 k-atoms each
    float E[MAX_SLOTS * 768];                 // proj[obj] + tq per object slot; compute_tq scratch between evaluations
    float x[2 * RT * 12];                     // inputs [RT][12]; reused as out[2][RT][12] once layer 1 has read them
    float tq[768];                            // single-stage t-branch (eval / PC kernels; the ODE keeps 6 in global)
    float b2[256];
    float times[8];
    double red[16];
    int obj[RT];
    int slot_base, nslots;
    unsigned long long full[NSTAGE], empty[NSTAGE], a_ready, d_full[2], d_free;
    uint32_t tmem_base;
};

// pipeline state carried across evaluations (every thread holds a copy, each role uses its own fields)
struct State {
    uint32_t loads = 0;      // producer: chunks issued so far
    uint32_t consumed = 0;   // MMA issuer: chunks consumed so far
    uint32_t a_phase = 0;    // MMA issuer: parity of the next a_ready completion
    uint32_t dfree_phase = 0;
    uint32_t dfull_phase[2] = {0, 0};  // epilogue: parity of the next d_full[i] completion
    float w1col[9];          // epilogue thread c: column c of the first pose-encoder layer
    float b1v;
    // cycle counters of one epilogue thread (phase breakdown of an evaluation, reported through `stats`)
    long long cyc_l1 = 0, cyc_wait1 = 0, cyc_epi1 = 0, cyc_waith = 0, cyc_epi2 = 0, cyc_fwd = 0;
};

__device__ __forceinline__ const uint8_t *chunk_src(const float *__restrict__ P, uint32_t q, int which) {
    return reinterpret_cast<const uint8_t *>(P + TrunkLayout::W_TC) + ((size_t)q * 2 + which) * IMG_BYTES;
}

template <int NPASS>
__device__ __forceinline__ void issue_chunk(Smem<NPASS> &S, const float *__restrict__ P, uint32_t L) {
    constexpr int NST = Smem<NPASS>::NSTAGE, IM = Smem<NPASS>::IMAGES;
    const uint32_t s = L % NST, q = L % NCHUNK;
    mbar_arrive_expect_tx(&S.full[s], IM * IMG_BYTES);
    for (int w = 0; w < IM; ++w) bulk_g2s(S.ring[s][w], chunk_src(P, q, w), IMG_BYTES, &S.full[s]);
}

// one-time setup / teardown (all threads call)
template <int NPASS>
__device__ __forceinline__ void setup(Smem<NPASS> &S, State &st, const float *__restrict__ P) {
    constexpr int NST = Smem<NPASS>::NSTAGE;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        mbar_init(&S.a_ready, 8);
        mbar_init(&S.d_full[0], 1);
        mbar_init(&S.d_full[1], 1);
        mbar_init(&S.d_free, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 256; i += NTHREADS) S.b2[i] = __ldg(P + TrunkLayout::B2 + i);
    if (tid >= 64) {  // epilogue thread c = tid - 64 owns column c of layer 1
        const int c = tid - 64;
#pragma unroll
        for (int k = 0; k < 9; ++k) st.w1col[k] = __ldg(P + TrunkLayout::W1T + k * 256 + c);
        st.b1v = __ldg(P + TrunkLayout::B1 + c);
    }
    __syncthreads();
    if (warp == 1) tmem_alloc(&S.tmem_base, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {  // prefill the ring
        for (uint32_t L = 0; L < (uint32_t)NST; ++L) issue_chunk<NPASS>(S, P, L);
        st.loads = NST;
    }
}

template <int NPASS>
__device__ __forceinline__ void teardown(Smem<NPASS> &S, State &st) {
    constexpr int NST = Smem<NPASS>::NSTAGE;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 1 && (tid & 31) == 0) {  // drain the NSTAGE chunks that are still in flight
        for (int i = 0; i < NST; ++i) {
            const uint32_t g = st.consumed + i;
            mbar_wait(&S.full[g % NST], (g / NST) & 1);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(S.tmem_base, TMEM_COLS);
}

// byte offset of (row, column n) inside an A buffer (4 k-atoms of [128 rows][64 k] bf16, SWIZZLE_128B)
__device__ __forceinline__ int a_offset(int row, int n) {
    return (n >> 6) * ATOM_BYTES + row * 128 + ((((n & 63) >> 3) ^ (row & 7)) << 4) + ((n & 7) << 1);
}

// f_theta for the 128 rows in S.x -> S.x (as out[0])[r*12 + c]; `tq` is this stage's t-branch (768 floats, any
// address space).  All 320 threads call; ends with __syncthreads().
template <int NPASS>
__device__ __forceinline__ void forward(const float *__restrict__ P, const float *__restrict__ proj, Smem<NPASS> &S,
                                        State &st, const float *tq) {
    constexpr int NST = Smem<NPASS>::NSTAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = S.tmem_base;
    const long long t_begin = clock64();

    // ---- (proj + tq) table for the objects this tile spans (rows of an object are contiguous) ----
    if (tid == 0) {
        int first = -1, last = -1;
        for (int r = 0; r < RT; ++r) {
            const int o = S.obj[r];
            if (o >= 0) { if (first < 0) first = o; last = o; }
        }
        S.slot_base = first < 0 ? 0 : first;
        S.nslots = first < 0 ? 0 : last - first + 1;
    }
    __syncthreads();
    const int nslots = S.nslots, slot_base = S.slot_base;
    const bool use_E = nslots <= MAX_SLOTS;
    if (use_E) {
        for (int i = tid; i < nslots * 192; i += NTHREADS) {
            const int s = i / 192, n4 = i - 192 * s;
            const float4 pj = __ldg(reinterpret_cast<const float4 *>(proj + (size_t)(slot_base + s) * 768) + n4);
            const float4 tv = *reinterpret_cast<const float4 *>(tq + 4 * n4);
            *reinterpret_cast<float4 *>(S.E + s * 768 + 4 * n4) = make_float4(pj.x + tv.x, pj.y + tv.y, pj.z + tv.z, pj.w + tv.w);
        }
    }
    __syncthreads();

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int i = 0; i < NCHUNK; ++i) {
                const uint32_t L = st.loads;
                mbar_wait(&S.empty[L % NST], ((L / NST) + 1) & 1);
                issue_chunk<NPASS>(S, P, L);
                st.loads = L + 1;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t a_hi = smem_u32(&S.abuf[0][0][0]);
            const uint32_t a_lo = smem_u32(&S.abuf[NPASS == 3 ? 1 : 0][0][0]);
            auto gemm = [&](uint32_t dcol) {
                for (int kc = 0; kc < 4; ++kc)
                    for (int nh = 0; nh < 2; ++nh) {
                        const uint32_t g = st.consumed;
                        const uint32_t s = g % NST;
                        mbar_wait(&S.full[s], (g / NST) & 1);
                        tc_fence_after();
                        const uint32_t b_hi = smem_u32(&S.ring[s][0][0]);
                        const uint32_t b_lo = smem_u32(&S.ring[s][NPASS == 3 ? 1 : 0][0]);
                        const uint32_t d = tmem + dcol + nh * 128;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t ao = kc * ATOM_BYTES + kk * 32, bo = kk * 32;
                            umma_bf16(d, make_desc(a_hi + ao), make_desc(b_hi + bo), kIdescN128, (kc | kk) ? 1u : 0u);
                            if (NPASS == 3) {
                                umma_bf16(d, make_desc(a_lo + ao), make_desc(b_hi + bo), kIdescN128, 1u);
                                umma_bf16(d, make_desc(a_hi + ao), make_desc(b_lo + bo), kIdescN128, 1u);
                            }
                        }
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
        const uint32_t A_hi = smem_u32(&S.abuf[0][0][0]);
        const uint32_t A_lo = smem_u32(&S.abuf[NPASS == 3 ? 1 : 0][0][0]);
        // Views of S.x / S.b2 / S.E derived from the dynamic shared array itself: the compiler then knows the
        // address space (LDS, freely schedulable) although S is only reachable through a generic pointer.
        extern __shared__ __align__(16) unsigned char gp_dyn_smem[];
        const uint32_t dyn0 = smem_u32(gp_dyn_smem);
        const float *sx = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.x) - dyn0));
        const float *sb2 = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.b2) - dyn0));
        const float *sE = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.E) - dyn0));
        // layer 1 (9 -> 256): this thread owns output column c for all 128 rows (weights in registers)
        long long t0 = clock64();
        {
            const int c = tid - 64;
            // (row-independent part of the swizzled offset of column c)
            const int c_atom = (c >> 6) * ATOM_BYTES, c_unit = (c & 63) >> 3, c_byte = (c & 7) << 1;
            for (int r0 = 0; r0 < RT; r0 += 4) {
                // four rows at a time: four independent FMA chains instead of one 9-deep dependent chain per row
                float4 xa[4], xb[4];
                float xc[4], a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    xa[u] = *reinterpret_cast<const float4 *>(sx + (r0 + u) * 12);
                    xb[u] = *reinterpret_cast<const float4 *>(sx + (r0 + u) * 12 + 4);
                    xc[u] = sx[(r0 + u) * 12 + 8];
                    a[u] = st.b1v;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xa[u].x, st.w1col[0], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xa[u].y, st.w1col[1], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xa[u].z, st.w1col[2], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xa[u].w, st.w1col[3], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xb[u].x, st.w1col[4], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xb[u].y, st.w1col[5], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xb[u].z, st.w1col[6], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaf(xb[u].w, st.w1col[7], a[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) a[u] = fmaxf(fmaf(xc[u], st.w1col[8], a[u]), 0.f);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int r = r0 + u;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(a[u]);
                    const int off = c_atom + r * 128 + ((c_unit ^ (r & 7)) << 4) + c_byte;
                    sts_u16(A_hi + off, __bfloat16_as_ushort(hi));
                    if (NPASS == 3) sts_u16(A_lo + off, __bfloat16_as_ushort(__float2bfloat16_rn(a[u] - __bfloat162float(hi))));
                }
            }
            fence_proxy_async();
            // every epilogue thread has now read the inputs: S.x may be reused as the output buffer
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) mbar_arrive(&S.a_ready);
        }
        // layer 2 epilogue: h2 = relu(D1 + b2) -> A buffers
        {
            long long t1 = clock64();
            st.cyc_l1 += t1 - t0;
            mbar_wait(&S.d_full[0], st.dfull_phase[0]); st.dfull_phase[0] ^= 1;
            tc_fence_after();
            t0 = clock64();
            st.cyc_wait1 += t0 - t1;
            for (int g = 0; g < 4; ++g) {
                uint32_t r[32];
                tmem_ld32(lane_addr + c0 + g * 32, r);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    const int n0 = c0 + g * 32 + j8 * 8;
                    float v[8];
#pragma unroll
                    const float4 ba = *reinterpret_cast<const float4 *>(sb2 + n0), bb = *reinterpret_cast<const float4 *>(sb2 + n0 + 4);
                    const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaxf(__uint_as_float(r[j8 * 8 + j]) + bv[j], 0.f);
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                    const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<const uint32_t *>(&p0); pk.y = *reinterpret_cast<const uint32_t *>(&p1);
                    pk.z = *reinterpret_cast<const uint32_t *>(&p2); pk.w = *reinterpret_cast<const uint32_t *>(&p3);
                    const int off = a_offset(row, n0);
                    sts_u4(A_hi + off, pk);
                    if (NPASS == 3) {
                        const float2 f0 = __bfloat1622float2(p0), f1 = __bfloat1622float2(p1);
                        const float2 f2 = __bfloat1622float2(p2), f3 = __bfloat1622float2(p3);
                        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v[0] - f0.x, v[1] - f0.y), l1 = __floats2bfloat162_rn(v[2] - f1.x, v[3] - f1.y);
                        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[4] - f2.x, v[5] - f2.y), l3 = __floats2bfloat162_rn(v[6] - f3.x, v[7] - f3.y);
                        uint4 pl;
                        pl.x = *reinterpret_cast<const uint32_t *>(&l0); pl.y = *reinterpret_cast<const uint32_t *>(&l1);
                        pl.z = *reinterpret_cast<const uint32_t *>(&l2); pl.w = *reinterpret_cast<const uint32_t *>(&l3);
                        sts_u4(A_lo + off, pl);
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.a_ready);
        }
        // heads: z = relu(D + proj + tq), out = z . Wo^T over this thread's 128 columns
        {
            const long long t1 = clock64();
            st.cyc_epi1 += t1 - t0;
            t0 = t1;
        }
        float acc[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) acc[c] = 0.f;
        const int o = S.obj[row];
        const float *erow = sE + (use_E && o >= 0 ? o - slot_base : 0) * 768;
        const float *prow = proj + (size_t)(o < 0 ? 0 : o) * 768;
#pragma unroll 1
        for (int h = 0; h < 3; ++h) {
            const int buf = (h == 1) ? 0 : 1;
            mbar_wait(&S.d_full[buf], st.dfull_phase[buf]); st.dfull_phase[buf] ^= 1;
            tc_fence_after();
            {
                const long long t1 = clock64();
                st.cyc_waith += t1 - t0;
                t0 = t1;
            }
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 1
            for (int g = 0; g < 4; ++g) {
                const int nb = h * 256 + c0 + g * 32;
                uint32_t r[32];
                tmem_ld32(lane_addr + (buf ? 256 : 0) + c0 + g * 32, r);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    // output-layer weights of 8 columns (L1 broadcast hits, independent of the accumulator)
                    float4 w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = *reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(nb + j8 * 8 + j) * 4);
                    float ev[8];
                    if (use_E) {
                        const float4 ea = *reinterpret_cast<const float4 *>(erow + nb + j8 * 8);
                        const float4 eb = *reinterpret_cast<const float4 *>(erow + nb + j8 * 8 + 4);
                        ev[0] = ea.x; ev[1] = ea.y; ev[2] = ea.z; ev[3] = ea.w; ev[4] = eb.x; ev[5] = eb.y; ev[6] = eb.z; ev[7] = eb.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) ev[j] = __ldg(prow + nb + j8 * 8 + j) + tq[nb + j8 * 8 + j];
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float z = fmaxf(__uint_as_float(r[j8 * 8 + j]) + ev[j], 0.f);
                        a0 = fmaf(z, w[j].x, a0); a1 = fmaf(z, w[j].y, a1); a2 = fmaf(z, w[j].z, a2);
                    }
                }
            }
            acc[h * 3 + 0] = a0; acc[h * 3 + 1] = a1; acc[h * 3 + 2] = a2;
            {
                const long long t1 = clock64();
                st.cyc_epi2 += t1 - t0;
                t0 = t1;
            }
            if (h == 0) {  // cols 256..511 may now be overwritten by head 2
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.d_free);
            }
        }
        tc_fence_before();
        float *out = S.x;  // inputs are dead since layer 1: reuse as out[2][RT][12]
#pragma unroll
        for (int c = 0; c < 9; ++c) out[half * RT * 12 + row * 12 + c] = acc[c];
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
        if (half == 0) {
#pragma unroll
            for (int c = 0; c < 9; ++c)
                out[row * 12 + c] = (out[row * 12 + c] + out[RT * 12 + row * 12 + c]) + __ldg(P + TrunkLayout::BO + c);
        }
    }
    __syncthreads();
    st.cyc_fwd += clock64() - t_begin;
}

}  // namespace tc
}  // namespace gp
