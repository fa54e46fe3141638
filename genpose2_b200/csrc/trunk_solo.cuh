// Tensor-core tile evaluator of the ScoreNet trunk, one-CTA-per-tile ("solo") edition: the throughput shape.
//
// The cluster evaluator (trunk_tc.cuh) spends four SMs on one 128-row tile to shorten the critical path of a small
// batch; every CTA of the cluster recomputes the pose encoder and keeps a private copy of the integrator state.
// With more tiles than the GPU has clusters that is the wrong trade: this evaluator gives every tile ONE CTA, no
// redundant work, no exchange, one copy of the state, and overlaps what is independent inside an evaluation -- the
// three 256 -> 256 head layers (three quarters of the multiply-adds) are issued back to back into alternating TMEM
// accumulators, so the epilogue of head h (TMEM -> + proj + tq -> ReLU -> 256 -> 3 output layer) runs under the
// MMAs of head h + 1.
//
// Per evaluation of a 128-row tile (one UMMA M = 128 accumulator, 512 TMEM columns):
//   D0[  0..255] = x  . W1^T (K = 16)        h1 = relu(D0 + b1) -> A buffer
//   D1[256..511] = h1 . W2^T                 h2 = relu(D1 + b2) -> A buffer (in place)
//   H0[  0..255] = h2 . Wh0^T   H1[256..511] = h2 . Wh1^T   H2[0..255] = h2 . Wh2^T (after H0 has been drained)
//   out[row][3h..3h+2] = relu(Hh + proj[obj] + tq) . Wo_h^T + bo
// Weights: 34 chunks per evaluation ([128 n][64 k] bf16 images, canonical K-major SWIZZLE_128B, 16 KB; hi and lo
// images alternate in split-bf16 mode), streamed from L2 by one producer lane with cp.async.bulk through an mbarrier
// ring of single images that runs ahead across layers and evaluations.
//
// 10 warps as in the cluster evaluator: warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue (row = 32 (warp % 4)
// + lane, column half = (warp - 2) / 4).  NPASS as in trunk_tc.cuh (1: bf16; 3: split-bf16, hi*hi + lo*hi + hi*lo).
#pragma once
#include "trunk_tc.cuh"

namespace gp {
namespace solo {

using namespace tc;

constexpr int NCHUNKS = 34;          // 2 (pose_encoder.0) + 8 (pose_encoder.2) + 3 x 8 (heads)
constexpr int SLOTS = 4;             // objects a tile may span for the shared-memory proj table

template <int NPASS>
struct Smem {
    static constexpr int IMAGES = NPASS == 3 ? 2 : 1;
    static constexpr int NSTAGE = NPASS == 3 ? 3 : 4;
    static constexpr int ENTRIES = NCHUNKS * IMAGES;   // ring entries (single images) per evaluation
    uint8_t ring[NSTAGE][IMG_BYTES];          // the struct sits on a 1024-byte boundary
    uint8_t abuf[IMAGES][4][ATOM_BYTES];      // x / h1 / h2 as A operand (hi, lo), 4 k-atoms each; scratch between evaluations
    float4 wo[768];                           // output-layer weights per head column (3 used)
    float tq[6 * 768];                        // t-branch, up to 6 stages
    float pj[SLOTS * 768];                    // proj rows of the objects of the tile
    float x[RT * XS];                         // inputs [RT][9]; overwritten with f_theta [RT][9] by forward()
    float b1[256], b2[256];
    float bo[12];
    float times[8];
    double red[16];
    unsigned long long full[NSTAGE], empty[NSTAGE], a_ready, dbar[5], hdrain;
    uint32_t tmem_base;
};

struct State {
    uint32_t loads = 0;      // producer: ring entries issued so far
    uint32_t consumed = 0;   // MMA issuer: ring entries consumed so far
    uint32_t a_phase = 0;    // MMA issuer: parity of the next a_ready completion
    uint32_t evals = 0;      // evaluations done (parity of the once-per-evaluation barriers)
    int tile_r0 = -1;        // tile whose proj table is loaded
    int slot_base = 0, nslots = 0, rpo = 1, nrows = 0;
    long long cyc_fwd = 0, cyc_l1 = 0, cyc_wait1 = 0, cyc_epi1 = 0, cyc_waith = 0, cyc_epi2 = 0, cyc_tail = 0;
};

// ring entry L of the endless per-CTA stream -> source image.  Chunks 0..9 are the pose-encoder chunks of the cluster
// layout; the head chunks [h][kc][nh] live in the W_SOLO region (trunk.cuh).
template <int NPASS>
__device__ __forceinline__ const uint8_t *entry_src(const float *__restrict__ P, uint32_t L) {
    constexpr int IM = Smem<NPASS>::IMAGES;
    const uint32_t e = L % (uint32_t)Smem<NPASS>::ENTRIES;
    const uint32_t q = e / IM, which = e % IM;
    const uint8_t *base = q < 10 ? reinterpret_cast<const uint8_t *>(P + TrunkLayout::W_TC) + (size_t)q * 2 * IMG_BYTES
                                 : reinterpret_cast<const uint8_t *>(P + TrunkLayout::W_SOLO) + (size_t)(q - 10) * 2 * IMG_BYTES;
    return base + (size_t)which * IMG_BYTES;
}

template <int NPASS>
__device__ __forceinline__ void issue_entry(Smem<NPASS> &S, const float *__restrict__ P, uint32_t L) {
    const uint32_t s = L % (uint32_t)Smem<NPASS>::NSTAGE;
    mbar_arrive_expect_tx(&S.full[s], IMG_BYTES);
    bulk_g2s(S.ring[s], entry_src<NPASS>(P, L), IMG_BYTES, &S.full[s]);
}

template <int NPASS>
__device__ __forceinline__ void setup(Smem<NPASS> &S, State &st, const float *__restrict__ P) {
    constexpr int NST = Smem<NPASS>::NSTAGE;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); }
        mbar_init(&S.a_ready, 8);
        for (int i = 0; i < 5; ++i) mbar_init(&S.dbar[i], 1);
        mbar_init(&S.hdrain, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 256; i += NTHREADS) {
        S.b1[i] = __ldg(P + TrunkLayout::B1 + i);
        S.b2[i] = __ldg(P + TrunkLayout::B2 + i);
    }
    for (int i = tid; i < 768; i += NTHREADS) S.wo[i] = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO) + i);
    if (tid < 12) S.bo[tid] = __ldg(P + TrunkLayout::BO + tid);
    __syncthreads();
    if (warp == 1) tmem_alloc(&S.tmem_base, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {  // prefill the ring
        for (uint32_t L = 0; L < (uint32_t)NST; ++L) issue_entry<NPASS>(S, P, L);
        st.loads = NST;
    }
    __syncwarp();
}

template <int NPASS>
__device__ __forceinline__ void teardown(Smem<NPASS> &S, State &st) {
    constexpr int NST = Smem<NPASS>::NSTAGE;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 1 && (tid & 31) == 0) {  // drain the entries that are still in flight
        for (int i = 0; i < NST; ++i) {
            const uint32_t g = st.consumed + i;
            mbar_wait(&S.full[g % NST], (g / NST) & 1);
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(S.tmem_base, TMEM_COLS);
}

template <int NPASS>
__device__ __forceinline__ void compute_tq_all(const float *__restrict__ P, Smem<NPASS> &S, int ns) {
    static_assert(sizeof(S.abuf) >= 2048 * sizeof(float), "compute_tq scratch");
    compute_tq(P, S.times, ns, reinterpret_cast<float *>(&S.abuf[0][0][0]), S.tq);
}

// rows r0.. of the batch become the current tile.  Callers synchronise the block (they fill S.x next) before forward().
template <int NPASS>
__device__ __forceinline__ void begin_tile(Smem<NPASS> &S, State &st, const float *__restrict__ proj, int r0, int N, int rpo) {
    if (st.tile_r0 == r0) return;
    st.tile_r0 = r0;
    st.rpo = rpo;
    st.nrows = max(0, min(RT, N - r0));
    const int tid = threadIdx.x;
    const int first = r0 / rpo, last = (min(r0 + RT, N) - 1) / rpo;
    st.slot_base = first;
    st.nslots = r0 < N ? last - first + 1 : 0;
    if (st.nslots <= SLOTS) {
        const float4 *src = reinterpret_cast<const float4 *>(proj + (size_t)first * 768);
        float4 *dst = reinterpret_cast<float4 *>(S.pj);
        for (int i = tid; i < st.nslots * 192; i += NTHREADS) dst[i] = __ldg(src + i);
    }
}

// f_theta for the 128 rows in S.x -> S.x [r*9 + c]; `tq` is this stage's t-branch (768 floats in shared memory).
// All 320 threads call; ends with __syncthreads().
template <int NPASS>
__device__ __noinline__ void forward(const float *__restrict__ P, const float *__restrict__ proj, Smem<NPASS> &S,
                                     State &st, const float *tq) {
    constexpr int NST = Smem<NPASS>::NSTAGE, EPE = Smem<NPASS>::ENTRIES;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t tmem = S.tmem_base;
    const uint32_t eph = st.evals & 1;
    const int r0 = st.tile_r0;
    const long long t_begin = clock64();

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int i = 0; i < EPE; ++i) {
                const uint32_t L = st.loads;
                mbar_wait(&S.empty[L % NST], ((L / NST) + 1) & 1);
                issue_entry<NPASS>(S, P, L);
                st.loads = L + 1;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t a_hi = smem_u32(&S.abuf[0][0][0]);
            const uint32_t a_lo = smem_u32(&S.abuf[NPASS == 3 ? 1 : 0][0][0]);
            // one weight chunk (hi image, then lo image in split mode) against k-atom `kc` of the A buffer
            // (`nk` 16-wide k steps) into accumulator columns d .. d + 127
            auto chunk = [&](uint32_t d, int kc, int nk, bool first) {
                {
                    const uint32_t g = st.consumed, s = g % NST;
                    mbar_wait(&S.full[s], (g / NST) & 1);
                    tc_fence_after();
                    const uint32_t b = smem_u32(&S.ring[s][0]);
                    for (int kk = 0; kk < nk; ++kk) {
                        const uint32_t ao = kc * ATOM_BYTES + kk * 32;
                        umma_bf16(d, make_desc(a_hi + ao), make_desc(b + kk * 32), kIdescN128, (first && kk == 0) ? 0u : 1u);
                        if (NPASS == 3) umma_bf16(d, make_desc(a_lo + ao), make_desc(b + kk * 32), kIdescN128, 1u);
                    }
                    umma_commit(&S.empty[s]);
                    st.consumed = g + 1;
                }
                if (NPASS == 3) {
                    const uint32_t g = st.consumed, s = g % NST;
                    mbar_wait(&S.full[s], (g / NST) & 1);
                    tc_fence_after();
                    const uint32_t b = smem_u32(&S.ring[s][0]);
                    for (int kk = 0; kk < nk; ++kk)
                        umma_bf16(d, make_desc(a_hi + kc * ATOM_BYTES + kk * 32), make_desc(b + kk * 32), kIdescN128, 1u);
                    umma_commit(&S.empty[s]);
                    st.consumed = g + 1;
                }
            };
            // pose_encoder.0: K = 16 (9 used) -> D0, cols 0..255
            mbar_wait(&S.a_ready, st.a_phase); st.a_phase ^= 1;
            tc_fence_after();
            chunk(tmem, 0, 1, true);
            chunk(tmem + 128, 0, 1, true);
            umma_commit(&S.dbar[0]);
            // pose_encoder.2 -> D1, cols 256..511
            mbar_wait(&S.a_ready, st.a_phase); st.a_phase ^= 1;
            tc_fence_after();
            for (int kc = 0; kc < 4; ++kc)
                for (int nh = 0; nh < 2; ++nh) chunk(tmem + 256 + nh * 128, kc, 4, kc == 0);
            umma_commit(&S.dbar[1]);
            // heads: H0 -> cols 0..255, H1 -> cols 256..511, H2 -> cols 0..255 once H0 has been read out
            mbar_wait(&S.a_ready, st.a_phase); st.a_phase ^= 1;
            tc_fence_after();
            for (int h = 0; h < 3; ++h) {
                if (h == 2) { mbar_wait(&S.hdrain, eph); tc_fence_after(); }
                const uint32_t d = tmem + (h & 1) * 256;
                for (int kc = 0; kc < 4; ++kc)
                    for (int nh = 0; nh < 2; ++nh) chunk(d + nh * 128, kc, 4, kc == 0);
                umma_commit(&S.dbar[2 + h]);
            }
        }
        __syncwarp();
    } else {
        // ---------------- epilogue warps ----------------
        const int e = warp - 2;
        const int row = 32 * (warp & 3) + lane;   // TMEM lane quarter is fixed by warp % 4
        const int half = e >> 2;                  // column half 0 / 1
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        const uint32_t A_hi = smem_u32(&S.abuf[0][0][0]);
        const uint32_t A_lo = smem_u32(&S.abuf[NPASS == 3 ? 1 : 0][0][0]);
        extern __shared__ __align__(16) unsigned char gp_dyn_smem[];
        const uint32_t dyn0 = smem_u32(gp_dyn_smem);
        const float *sx = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.x) - dyn0));
        const float *sb1 = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.b1) - dyn0));
        const float *sb2 = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.b2) - dyn0));
        const float *spj = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.pj) - dyn0));
        const float *stq = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(tq) - dyn0));
        const float4 *swo = reinterpret_cast<const float4 *>(gp_dyn_smem + (smem_u32(S.wo) - dyn0));
        float *sxw = reinterpret_cast<float *>(gp_dyn_smem + (smem_u32(S.x) - dyn0));
        float *scratch = reinterpret_cast<float *>(gp_dyn_smem + (smem_u32(&S.abuf[0][0][0]) - dyn0));

        long long t0 = clock64();
        // inputs -> A operand: k 0..8 of atom 0 (k 9..15 zero), one row per thread of the first four epilogue warps
        if (half == 0) {
            float xv[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) xv[c] = sx[row * XS + c];
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const float v0 = 2 * p < 9 ? xv[2 * p] : 0.f, v1 = 2 * p + 1 < 9 ? xv[2 * p + 1] : 0.f;
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                const float2 hf = __bfloat1622float2(h2);
                const __nv_bfloat162 l2 = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
                hi[p] = *reinterpret_cast<const uint32_t *>(&h2);
                lo[p] = *reinterpret_cast<const uint32_t *>(&l2);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int off = row * 128 + ((j ^ (row & 7)) << 4);
                sts_u4(A_hi + off, make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]));
                if (NPASS == 3) sts_u4(A_lo + off, make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]));
            }
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a_ready);

        // h1 = relu(D0 + b1) -> A buffers
        mbar_wait(&S.dbar[0], eph);
        tc_fence_after();
        epi_hidden<NPASS>(lane_addr, sb1, row, half * 128, A_hi, A_lo);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a_ready);
        { const long long t1 = clock64(); st.cyc_l1 += t1 - t0; t0 = t1; }
        // h2 = relu(D1 + b2) -> A buffers
        mbar_wait(&S.dbar[1], eph);
        tc_fence_after();
        { const long long t1 = clock64(); st.cyc_wait1 += t1 - t0; t0 = t1; }
        epi_hidden<NPASS>(lane_addr + 256, sb2, row, half * 128, A_hi, A_lo);
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&S.a_ready);
        { const long long t1 = clock64(); st.cyc_epi1 += t1 - t0; t0 = t1; }

        // heads: z = relu(H + proj + tq), out = z . Wo^T over this thread's 128 columns of each head
        const int o = row < st.nrows ? (r0 + row) / st.rpo : -1;
        const bool use_pj = st.nslots <= SLOTS;
        const float *pjrow = spj + (use_pj && o >= 0 ? o - st.slot_base : 0) * 768;
        const float *prow = proj + (size_t)(o < 0 ? 0 : o) * 768;
        float acc[9];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            mbar_wait(&S.dbar[2 + h], eph);
            tc_fence_after();
            { const long long t1 = clock64(); st.cyc_waith += t1 - t0; t0 = t1; }
            const int cb = h * 256 + half * 128;   // first head column (0..767) of this thread
            const uint32_t taddr = lane_addr + (h & 1) * 256 + half * 128;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
            uint32_t r[2][32];
            tmem_ld32_nowait(taddr, r[0]);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g + 1 < 4) tmem_ld32_nowait(taddr + (g + 1) * 32, r[(g + 1) & 1]);
                tmem_ld_wait();
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    const int c0 = cb + g * 32 + j8 * 8;
                    float4 w[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = swo[c0 + j];
                    const float4 ta = *reinterpret_cast<const float4 *>(stq + c0);
                    const float4 tb = *reinterpret_cast<const float4 *>(stq + c0 + 4);
                    const float *ebase = use_pj ? pjrow + c0 : prow + c0;
                    const float4 ea = *reinterpret_cast<const float4 *>(ebase);
                    const float4 eb = *reinterpret_cast<const float4 *>(ebase + 4);
                    const float ev[8] = {ea.x + ta.x, ea.y + ta.y, ea.z + ta.z, ea.w + ta.w,
                                         eb.x + tb.x, eb.y + tb.y, eb.z + tb.z, eb.w + tb.w};
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float z = fmaxf(__uint_as_float(r[g & 1][j8 * 8 + j]) + ev[j], 0.f);
                        a0 = fmaf(z, w[j].x, a0); a1 = fmaf(z, w[j].y, a1); a2 = fmaf(z, w[j].z, a2);
                    }
                }
            }
            acc[h * 3 + 0] = a0; acc[h * 3 + 1] = a1; acc[h * 3 + 2] = a2;
            if (h == 0) {   // H0's accumulator columns have been read out: the MMA warp may overwrite them with H2
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.hdrain);
            }
            { const long long t1 = clock64(); st.cyc_epi2 += t1 - t0; t0 = t1; }
        }
        tc_fence_before();
        // the two column halves of a row meet through the (now idle) A buffer; half 0 writes the output tile
        if (half == 1) {
#pragma unroll
            for (int c = 0; c < 9; ++c) scratch[c * RT + row] = acc[c];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
        if (half == 0) {
            const float *sbo = reinterpret_cast<const float *>(gp_dyn_smem + (smem_u32(S.bo) - dyn0));
#pragma unroll
            for (int c = 0; c < 9; ++c) sxw[row * XS + c] = (acc[c] + scratch[c * RT + row]) + sbo[c];
        }
        st.cyc_tail += clock64() - t0;
    }
    __syncthreads();
    st.evals += 1;
    st.cyc_fwd += clock64() - t_begin;
}

}  // namespace solo
}  // namespace gp
