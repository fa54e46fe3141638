// Packed layout of a PoseScoreNet / PoseEnergyNet trunk and the FP32 (FFMA) tile evaluator shared
// by the ODE / PC / energy / single-eval kernels.
//
// Hoisting: the three head inputs are cat[pts_feat(1024) | t_feat(128) |
// pose_feat(256)] (scorenet.py:249).  The pts_feat columns are constant per object -> `proj`
// (gp_trunk_project, once per call); the t_feat columns are constant per RHS evaluation -> `tq`
// (compute_tq, once per stage time, shared by every row); only the pose_feat columns
// (256 -> 768) plus the pose encoder (9 -> 256 -> 256) and the 256 -> 3 output layers are
// evaluated per row: 266 752 MAC instead of 1 167 872.
#pragma once
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gp {

// offsets in floats inside the packed blob
struct TrunkLayout {
    static constexpr size_t W1T = 0;                           // [9][256]    k-major
    static constexpr size_t B1 = W1T + 9 * 256;                // [256]
    static constexpr size_t W2T = B1 + 256;                    // [256][256]  k-major
    static constexpr size_t B2 = W2T + 256 * 256;              // [256]
    static constexpr size_t FOUR = B2 + 256;                   // [64]
    static constexpr size_t WTT = FOUR + 64;                   // [128][128]  k-major
    static constexpr size_t BT = WTT + 128 * 128;              // [128]
    static constexpr size_t WHP = BT + 128;                    // [3 heads][256 k][256 n] k-major, pose_feat cols
    static constexpr size_t WHT = WHP + 256 * 768;             // [128][768]  k-major, t_feat cols
    static constexpr size_t WHF = WHT + 128 * 768;             // [768][1024] n-major, pts_feat cols
    static constexpr size_t BH = WHF + 768 * 1024;             // [768]
    static constexpr size_t WO = BH + 768;                     // [768][4]    (3 used) out weights
    static constexpr size_t BO = WO + 768 * 4;                 // [12]        (9 used)
    static constexpr size_t F32_END = BO + 12;
    // bf16 B operands for the tensor-core path (tcgen05): 34 chunks x {hi, lo} ready-to-copy shared-memory
    // images of [128 rows][64 k] bf16 in the canonical K-major SWIZZLE_128B layout (16 KB each):
    //   q = 0, 1          pose_encoder.0, n-half q (k < 9 populated)
    //   q = 2 + 2*kc + nh pose_encoder.2, k-atom kc, n-half nh
    //   q = 10 + 6*r + 2*h + j   head h, the 64 output columns owned by cluster rank r, rows 0..63 = k-atom 2j,
    //                            rows 64..127 = k-atom 2j+1 (pose_feat columns of the first head layer)
    // lo = bf16(w - hi) for the split-bf16 (fp32-class) mode.  Offsets counted in floats; 1024-byte aligned.
    static constexpr size_t TC_CHUNKS = 34;
    static constexpr size_t W_TC = (F32_END + 255) / 256 * 256;
    // the one-CTA-per-tile evaluator (trunk_solo.cuh) multiplies whole heads: 24 more chunks x {hi, lo},
    //   q' = 8*h + 2*kc + nh   head h, k-atom kc, output columns 128*nh .. 128*nh + 127 (pose_feat columns)
    static constexpr size_t SOLO_CHUNKS = 24;
    static constexpr size_t W_SOLO = W_TC + TC_CHUNKS * 2 * (128 * 64 / 2);
    // the cluster evaluator multiplies a rank's three 64-column head slices as one N = 192 MMA: per (rank r, k-atom kc)
    // one image [192 n][64 k] x {hi, lo} (24 KB each), row 64*h + c = head h, output column 64*r + c
    static constexpr size_t WIDE_IMG_FLOATS = 192 * 64 / 2;
    static constexpr size_t W_WIDE = W_SOLO + SOLO_CHUNKS * 2 * (128 * 64 / 2);
    static constexpr size_t END = W_WIDE + 4 * 4 * 2 * WIDE_IMG_FLOATS;
};

constexpr float kFloatPi = 3.14159265358979323846f;  // np.pi cast to float32 by torch
constexpr int HS = 260;                              // padded row stride of activations in smem

// sigma(t) as the network sees it: float32 tensor math `0.01 * 5000 ** t` (sde.py:15-18 via
// scorenet.py:250).  Computed through double and rounded once (<= 0.5 ulp of the exact value).
__device__ __forceinline__ float sigma_f32(float t) {
    return 0.01f * (float)pow(5000.0, (double)t);
}
// g(t)^2 in float64 as samplers.py:213-219 computes it under numpy 2: sigma in f64, the sqrt
// constant is a float32 tensor (sde.py:24-26) promoted to f64.
__device__ __forceinline__ double diffusion_f64(double t) {
    const double sigma = 0.01 * pow(5000.0, t);
    const float c = sqrtf((float)(2.0 * (log(50.0) - log(0.01))));
    return sigma * (double)c;
}
__device__ __forceinline__ float diffusion_f32(float t) {
    const float c = sqrtf((float)(2.0 * (log(50.0) - log(0.01))));
    return sigma_f32(t) * c;
}

// --------------------------------------------------------------------------------------------
// t-branch: tq[s][n] = Wh[:, 1024:1152] @ relu(Wt @ [sin, cos](t_s * W * 2pi) + bt), n < 768.
// Whole block cooperates; ns <= 6 stage times.
// --------------------------------------------------------------------------------------------
// `ncols` head columns are produced; local column i is global column colmap(i) (a CTA of the tensor-core
// cluster evaluator only needs the 3 x 64 columns it owns).  s_tq is [ns][ncols]; `scratch` holds 2048 floats.
// Runs once per step between evaluations, so it is written for latency: every thread keeps 32 independent
// weight loads in flight, the stage values of one k sit in one 32-byte shared-memory row (two broadcast loads
// feed six FMAs), and the loops stay rolled so the code is fetched once.
template <class ColMap>
__device__ __noinline__ void compute_tq_cols(const float *__restrict__ P, const float *s_times, int ns,
                                             float *scratch, float *s_tq, int ncols, ColMap colmap) {
    const int tid = threadIdx.x, nt = blockDim.x;
    float *four8 = scratch, *tfeat8 = scratch + 1024;  // [128 k][8 stages]
    const uint32_t four_a = tc::smem_u32(four8), tfeat_a = tc::smem_u32(tfeat8);
    for (int i = tid; i < 8 * 64; i += nt) {
        const int s = i >> 6, j = i & 63;
        float sn = 0.f, cs = 0.f;
        if (s < ns) {
            // x[:, None] * W[None, :] * 2 * np.pi, left to right in float32 (scorenet.py:87)
            const float arg = __fmul_rn(__fmul_rn(__fmul_rn(s_times[s], __ldg(P + TrunkLayout::FOUR + j)), 2.0f), kFloatPi);
            sincosf(arg, &sn, &cs);
        }
        four8[j * 8 + s] = sn;
        four8[(64 + j) * 8 + s] = cs;
    }
    __syncthreads();
    // t_feat = relu(Wt . [sin, cos] + bt): thread pair (2j, 2j+1) splits k, all stages at once
    if (tid < 256) {
        const int j = tid >> 1, kp = tid & 1;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float *w = P + TrunkLayout::WTT + j + (size_t)(64 * kp) * 128;
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
            float wv[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) wv[u] = __ldg(w + (b * 32 + u) * 128);
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const uint32_t fa = four_a + (uint32_t)(64 * kp + b * 32 + u) * 32;
                const float4 f0 = tc::lds_f4_sync(fa);
                const float2 f1 = tc::lds_f2_sync(fa + 16);
                acc[0] = fmaf(f0.x, wv[u], acc[0]); acc[1] = fmaf(f0.y, wv[u], acc[1]); acc[2] = fmaf(f0.z, wv[u], acc[2]);
                acc[3] = fmaf(f0.w, wv[u], acc[3]); acc[4] = fmaf(f1.x, wv[u], acc[4]); acc[5] = fmaf(f1.y, wv[u], acc[5]);
            }
        }
        const float bt = __ldg(P + TrunkLayout::BT + j);
#pragma unroll
        for (int s = 0; s < 6; ++s) {
            const float v = acc[s] + __shfl_xor_sync(0xffffffffu, acc[s], 1);
            if (kp == 0) tfeat8[j * 8 + s] = fmaxf(v + bt, 0.f);
        }
    }
    __syncthreads();
    for (int n = tid; n < ncols; n += nt) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float *w = P + TrunkLayout::WHT + colmap(n);
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
            float wv[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) wv[u] = __ldg(w + (b * 32 + u) * 768);
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const uint32_t fa = tfeat_a + (uint32_t)(b * 32 + u) * 32;
                const float4 f0 = tc::lds_f4_sync(fa);
                const float2 f1 = tc::lds_f2_sync(fa + 16);
                acc[0] = fmaf(f0.x, wv[u], acc[0]); acc[1] = fmaf(f0.y, wv[u], acc[1]); acc[2] = fmaf(f0.z, wv[u], acc[2]);
                acc[3] = fmaf(f0.w, wv[u], acc[3]); acc[4] = fmaf(f1.x, wv[u], acc[4]); acc[5] = fmaf(f1.y, wv[u], acc[5]);
            }
        }
#pragma unroll
        for (int s = 0; s < 6; ++s)
            if (s < ns) s_tq[s * ncols + n] = acc[s];
    }
    __syncthreads();
}
__device__ __forceinline__ void compute_tq(const float *__restrict__ P, const float *s_times, int ns,
                                           float *scratch, float *s_tq) {
    compute_tq_cols(P, s_times, ns, scratch, s_tq, 768, [](int n) { return n; });
}

// --------------------------------------------------------------------------------------------
// FP32 tile evaluator: RT = 4*RPT rows per CTA, 8 compute warps + 1 TMA producer warp.
//   compute thread (ty = tid>>6, tx = tid&63) owns rows ty*RPT..+RPT-1 and columns 4*tx..4*tx+3 of each
//   256-wide output chunk.  The weights (1 MB of fp32 per evaluation, L2 resident) are streamed ONCE per
//   CTA through a shared-memory ring of 32 KB chunks ([32 k][256 n]) with cp.async.bulk + mbarrier by the
//   producer warp, which runs ahead across layers and evaluations; activations are broadcast from
//   shared memory; every inner-loop operand comes from shared memory, so the loop is FFMA bound.
// --------------------------------------------------------------------------------------------
constexpr int W_NS = 3;                      // ring stages
constexpr int W_CHUNK_FLOATS = 32 * 256;     // [32 k][256 n] fp32 = 32 KB
constexpr int W_NCHUNK = 32;                 // 8 (pose_encoder.2) + 3 heads x 8
constexpr int SIMT_COMPUTE_THREADS = 256;
constexpr int SIMT_THREADS = 288;

template <int RPT>
struct TileSmem {
    static constexpr int RT = 4 * RPT;
    float ring[W_NS][W_CHUNK_FLOATS];   // keep first: 16-byte aligned destination of the bulk copies
    float h1[RT * HS];
    float h2[RT * HS];
    float x[RT * 12];       // input poses, padded rows
    float out[2][RT * 12];  // per-half-warp-pair partial head outputs
    int obj[RT];            // object index of each row (-1: padding row)
    float tq[6 * 768];      // t-branch of up to 6 stage times
    float tqs[2048];        // compute_tq scratch
    float times[8];
    double red[16];
    unsigned long long full[W_NS], empty[W_NS];
};

struct SimtCtx {
    float w1col[9];
    float b1v;
    uint32_t loads = 0;      // producer lane: bulk copies issued
    uint32_t consumed = 0;   // compute threads: chunks consumed
};

__device__ __forceinline__ const float *weight_chunk(const float *__restrict__ P, uint32_t q) {
    return q < 8 ? P + TrunkLayout::W2T + (size_t)q * W_CHUNK_FLOATS
                 : P + TrunkLayout::WHP + (size_t)(q - 8) * W_CHUNK_FLOATS;  // heads are stored [h][k][n]
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// acc[i][0..3] += sum_k h[row0+i][k] * W[k][col..col+3] for the next 8 ring chunks (k = 0..255)
template <int RPT, class SM>
__device__ __forceinline__ void gemm256_ring(SM &sm, SimtCtx &ctx, const float *s_h, int row0, int col,
                                             float (&acc)[RPT][4]) {
    const int lane = threadIdx.x & 31;
    for (int c = 0; c < 8; ++c) {
        const uint32_t g = ctx.consumed;
        const uint32_t st = g % W_NS;
        tc::mbar_wait(&sm.full[st], (g / W_NS) & 1);
        const float *w = &sm.ring[st][col];
        const float *hrow = s_h + row0 * HS + c * 32;
#pragma unroll 2
        for (int k = 0; k < 32; k += 4) {
            const float4 w0 = *reinterpret_cast<const float4 *>(w + (k + 0) * 256);
            const float4 w1 = *reinterpret_cast<const float4 *>(w + (k + 1) * 256);
            const float4 w2 = *reinterpret_cast<const float4 *>(w + (k + 2) * 256);
            const float4 w3 = *reinterpret_cast<const float4 *>(w + (k + 3) * 256);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float4 a = *reinterpret_cast<const float4 *>(hrow + i * HS + k);
                acc[i][0] = fmaf(a.x, w0.x, acc[i][0]);
                acc[i][1] = fmaf(a.x, w0.y, acc[i][1]);
                acc[i][2] = fmaf(a.x, w0.z, acc[i][2]);
                acc[i][3] = fmaf(a.x, w0.w, acc[i][3]);
                acc[i][0] = fmaf(a.y, w1.x, acc[i][0]);
                acc[i][1] = fmaf(a.y, w1.y, acc[i][1]);
                acc[i][2] = fmaf(a.y, w1.z, acc[i][2]);
                acc[i][3] = fmaf(a.y, w1.w, acc[i][3]);
                acc[i][0] = fmaf(a.z, w2.x, acc[i][0]);
                acc[i][1] = fmaf(a.z, w2.y, acc[i][1]);
                acc[i][2] = fmaf(a.z, w2.z, acc[i][2]);
                acc[i][3] = fmaf(a.z, w2.w, acc[i][3]);
                acc[i][0] = fmaf(a.w, w3.x, acc[i][0]);
                acc[i][1] = fmaf(a.w, w3.y, acc[i][1]);
                acc[i][2] = fmaf(a.w, w3.z, acc[i][2]);
                acc[i][3] = fmaf(a.w, w3.w, acc[i][3]);
            }
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&sm.empty[st]);
        ctx.consumed = g + 1;
    }
}

template <class SM>
__device__ __forceinline__ void simt_setup(SM &sm, SimtCtx &ctx, const float *__restrict__ P) {
    const int tid = threadIdx.x;
    if (tid < SIMT_COMPUTE_THREADS) {
#pragma unroll
        for (int k = 0; k < 9; ++k) ctx.w1col[k] = __ldg(P + TrunkLayout::W1T + k * 256 + tid);
        ctx.b1v = __ldg(P + TrunkLayout::B1 + tid);
    }
    if (tid == 0) {
        for (int s = 0; s < W_NS; ++s) { tc::mbar_init(&sm.full[s], 1); tc::mbar_init(&sm.empty[s], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == SIMT_COMPUTE_THREADS) {  // producer lane: prefill the ring
        for (int s = 0; s < W_NS; ++s) {
            tc::mbar_arrive_expect_tx(&sm.full[s], W_CHUNK_FLOATS * 4);
            tc::bulk_g2s(sm.ring[s], weight_chunk(P, s), W_CHUNK_FLOATS * 4, &sm.full[s]);
        }
        ctx.loads = W_NS;
    }
}

template <class SM>
__device__ __forceinline__ void simt_teardown(SM &sm, SimtCtx &ctx) {
    if (threadIdx.x == 0) {  // the W_NS prefetched chunks are still in flight: let them land before exit
        for (int i = 0; i < W_NS; ++i) {
            const uint32_t g = ctx.consumed + i;
            tc::mbar_wait(&sm.full[g % W_NS], (g / W_NS) & 1);
        }
    }
    __syncthreads();
}

// Evaluates f_theta (un-normalised head outputs, 9 per row) for the RT rows whose inputs sit in sm.x;
// result lands in sm.out[0][r*12 + c] (c < 9).  `s_tq` is this stage's t-branch.  All 288 threads call;
// ends with a __syncthreads().
template <int RPT>
__device__ __forceinline__ void tile_forward(const float *__restrict__ P, const float *__restrict__ proj,
                                             TileSmem<RPT> &sm, SimtCtx &ctx, const float *s_tq) {
    constexpr int RT = 4 * RPT;
    const int tid = threadIdx.x;
    if (tid >= SIMT_COMPUTE_THREADS) {
        // ---------------- TMA producer: 32 chunks per evaluation, running W_NS chunks ahead ----------------
        if (tid == SIMT_COMPUTE_THREADS) {
            for (int i = 0; i < W_NCHUNK; ++i) {
                const uint32_t L = ctx.loads;
                const uint32_t st = L % W_NS;
                tc::mbar_wait(&sm.empty[st], ((L / W_NS) + 1) & 1);
                tc::mbar_arrive_expect_tx(&sm.full[st], W_CHUNK_FLOATS * 4);
                tc::bulk_g2s(sm.ring[st], weight_chunk(P, L % W_NCHUNK), W_CHUNK_FLOATS * 4, &sm.full[st]);
                ctx.loads = L + 1;
            }
        }
        __syncwarp();
    } else {
        const int tx = tid & 63, ty = tid >> 6;
        const int row0 = ty * RPT, col = 4 * tx;

        // layer 1: thread n = tid computes column n for every row
        for (int r = 0; r < RT; ++r) {
            const float4 xa = *reinterpret_cast<const float4 *>(sm.x + r * 12);
            const float4 xb = *reinterpret_cast<const float4 *>(sm.x + r * 12 + 4);
            const float xc = sm.x[r * 12 + 8];
            float a = ctx.b1v;
            a = fmaf(xa.x, ctx.w1col[0], a); a = fmaf(xa.y, ctx.w1col[1], a); a = fmaf(xa.z, ctx.w1col[2], a);
            a = fmaf(xa.w, ctx.w1col[3], a); a = fmaf(xb.x, ctx.w1col[4], a); a = fmaf(xb.y, ctx.w1col[5], a);
            a = fmaf(xb.z, ctx.w1col[6], a); a = fmaf(xb.w, ctx.w1col[7], a); a = fmaf(xc, ctx.w1col[8], a);
            sm.h1[r * HS + tid] = fmaxf(a, 0.f);
        }
        bar_compute();

        // layer 2: h2 = relu(h1 @ W2^T + b2)
        {
            float acc[RPT][4];
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::B2 + col));
#pragma unroll
            for (int i = 0; i < RPT; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
            gemm256_ring<RPT>(sm, ctx, sm.h1, row0, col, acc);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                float4 v = make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
                *reinterpret_cast<float4 *>(sm.h2 + (row0 + i) * HS + col) = v;
            }
        }
        bar_compute();

        // heads: z = relu(h2 @ Whp^T + proj[obj] + tq);  out = z @ Wo^T
        const int lane = tid & 31, half = (tid >> 5) & 1;
#pragma unroll 1
        for (int h = 0; h < 3; ++h) {
            float acc[RPT][4];
            const float4 tq = *reinterpret_cast<const float4 *>(s_tq + h * 256 + col);
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const int o = sm.obj[row0 + i];
                float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
                if (o >= 0) pj = __ldg(reinterpret_cast<const float4 *>(proj + (size_t)o * 768 + h * 256 + col));
                acc[i][0] = pj.x + tq.x; acc[i][1] = pj.y + tq.y; acc[i][2] = pj.z + tq.z; acc[i][3] = pj.w + tq.w;
            }
            gemm256_ring<RPT>(sm, ctx, sm.h2, row0, col, acc);
            // output layer of this head: 3 dot products over this thread's 4 columns, then the 64 column-threads
            // of the row group are reduced: 32 lanes by shuffle, the 2 warps through shared memory
            const float4 wo0 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 0) * 4));
            const float4 wo1 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 1) * 4));
            const float4 wo2 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 2) * 4));
            const float4 wo3 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 3) * 4));
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const float z0 = fmaxf(acc[i][0], 0.f), z1 = fmaxf(acc[i][1], 0.f);
                const float z2 = fmaxf(acc[i][2], 0.f), z3 = fmaxf(acc[i][3], 0.f);
                float p3[3];
                p3[0] = fmaf(z3, wo3.x, fmaf(z2, wo2.x, fmaf(z1, wo1.x, z0 * wo0.x)));
                p3[1] = fmaf(z3, wo3.y, fmaf(z2, wo2.y, fmaf(z1, wo1.y, z0 * wo0.y)));
                p3[2] = fmaf(z3, wo3.z, fmaf(z2, wo2.z, fmaf(z1, wo1.z, z0 * wo0.z)));
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float v = p3[c];
                    v += __shfl_xor_sync(0xffffffffu, v, 16);
                    v += __shfl_xor_sync(0xffffffffu, v, 8);
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    if (lane == 0) sm.out[half][(row0 + i) * 12 + h * 3 + c] = v;
                }
            }
        }
        bar_compute();
        for (int i = tid; i < RT * 9; i += SIMT_COMPUTE_THREADS) {
            const int r = i / 9, c = i - 9 * r;
            sm.out[0][r * 12 + c] = (sm.out[0][r * 12 + c] + sm.out[1][r * 12 + c]) + __ldg(P + TrunkLayout::BO + c);
        }
    }
    __syncthreads();
}

// deterministic block-wide sum of one double per thread (<= 16 warps); result valid in all threads
__device__ __forceinline__ double block_sum(double v, double *s_red /*[16]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    return t;
}

// Gram-Schmidt exactly as normalize_rotation/rotation_6d_to_matrix (misc.py:327-344,
// rotation_conversions.py:556-577; F.normalize eps = 1e-12): writes b1 -> v[0:3], b2 -> v[3:6]
template <typename T>
__device__ __forceinline__ void gram_schmidt6(T *v) {
    const T eps = (T)1e-12;
    T n1 = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    n1 = n1 > eps ? n1 : eps;
    const T b1x = v[0] / n1, b1y = v[1] / n1, b1z = v[2] / n1;
    const T d = b1x * v[3] + b1y * v[4] + b1z * v[5];
    T cx = v[3] - d * b1x, cy = v[4] - d * b1y, cz = v[5] - d * b1z;
    T n2 = sqrt(cx * cx + cy * cy + cz * cz);
    n2 = n2 > eps ? n2 : eps;
    v[0] = b1x; v[1] = b1y; v[2] = b1z;
    v[3] = cx / n2; v[4] = cy / n2; v[5] = cz / n2;
}

}  // namespace gp
