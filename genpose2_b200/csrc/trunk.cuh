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
    static constexpr size_t WHP = BT + 128;                    // [256][768]  k-major, pose_feat cols
    static constexpr size_t WHT = WHP + 256 * 768;             // [128][768]  k-major, t_feat cols
    static constexpr size_t WHF = WHT + 128 * 768;             // [768][1024] n-major, pts_feat cols
    static constexpr size_t BH = WHF + 768 * 1024;             // [768]
    static constexpr size_t WO = BH + 768;                     // [768][4]    (3 used) out weights
    static constexpr size_t BO = WO + 768 * 4;                 // [12]        (9 used)
    static constexpr size_t F32_END = BO + 12;
    // bf16 B operands for the tensor-core path (tcgen05), as 16 ready-to-copy shared-memory images of
    // [256 n][64 k] bf16 in the canonical K-major SWIZZLE_128B layout (32 KB each): chunks 0-3 = W2
    // k-atoms, chunks 4+4h+a = head h k-atom a.  Offsets counted in floats; 1024-byte aligned.
    static constexpr size_t W_TC = (F32_END + 255) / 256 * 256;
    static constexpr size_t END = W_TC + 16 * (256 * 64 / 2);
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
// Whole block cooperates; ns <= 6 stage times; s_four/s_tfeat are [6][128] scratch.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void compute_tq(const float *__restrict__ P, const float *s_times, int ns,
                                           float *s_four, float *s_tfeat, float *s_tq) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < ns * 64; i += nt) {
        const int s = i >> 6, j = i & 63;
        // x[:, None] * W[None, :] * 2 * np.pi, left to right in float32 (scorenet.py:87)
        const float arg = __fmul_rn(__fmul_rn(__fmul_rn(s_times[s], __ldg(P + TrunkLayout::FOUR + j)), 2.0f), kFloatPi);
        float sn, cs;
        sincosf(arg, &sn, &cs);
        s_four[s * 128 + j] = sn;
        s_four[s * 128 + 64 + j] = cs;
    }
    __syncthreads();
    for (int i = tid; i < ns * 128; i += nt) {
        const int s = i >> 7, j = i & 127;
        float acc = __ldg(P + TrunkLayout::BT + j);
        const float *w = P + TrunkLayout::WTT + j;
        const float *f = s_four + s * 128;
#pragma unroll 8
        for (int k = 0; k < 128; ++k) acc = fmaf(f[k], __ldg(w + k * 128), acc);
        s_tfeat[i] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int n = tid; n < 768; n += nt) {
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float *w = P + TrunkLayout::WHT + n;
#pragma unroll 4
        for (int k = 0; k < 128; ++k) {
            const float wv = __ldg(w + k * 768);
#pragma unroll
            for (int s = 0; s < 6; ++s)
                if (s < ns) acc[s] = fmaf(s_tfeat[s * 128 + k], wv, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < 6; ++s)
            if (s < ns) s_tq[s * 768 + n] = acc[s];
    }
    __syncthreads();
}

// --------------------------------------------------------------------------------------------
// FP32 tile evaluator: RT = 4*RPT rows per CTA of 256 threads.
//   thread (ty = tid>>6, tx = tid&63) owns rows ty*RPT..+RPT-1 and columns 4*tx..4*tx+3 of each
//   256-wide output chunk; weights stream from L2/L1 as float4, activations broadcast from smem.
// --------------------------------------------------------------------------------------------
template <int RPT>
struct TileSmem {
    static constexpr int RT = 4 * RPT;
    float h1[RT * HS];
    float h2[RT * HS];
    float x[RT * 12];       // input poses, padded rows
    float out[2][RT * 12];  // per-half-warp-pair partial head outputs
    int obj[RT];            // object index of each row (-1: padding row)
    float tq[6 * 768];      // t-branch of up to 6 stage times
    float four[6 * 128];
    float tfeat[6 * 128];
    float times[8];
    double red[16];
    float trow[RT];
};

template <int RPT>
__device__ __forceinline__ void gemm256(const float *__restrict__ Wt, int ldw, const float *s_h,
                                        int row0, float (&acc)[RPT][4]) {
    // weights for k..k+3 are fetched one iteration ahead (register double buffering): the L2/L1
    // latency of the streamed weight rows is the dominant stall of this loop otherwise
    float4 w0 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)0 * ldw));
    float4 w1 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)1 * ldw));
    float4 w2 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)2 * ldw));
    float4 w3 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)3 * ldw));
#pragma unroll 2
    for (int k = 0; k < 256; k += 4) {
        const int kn = (k + 4 < 256) ? k + 4 : k;
        const float4 n0 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(kn + 0) * ldw));
        const float4 n1 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(kn + 1) * ldw));
        const float4 n2 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(kn + 2) * ldw));
        const float4 n3 = __ldg(reinterpret_cast<const float4 *>(Wt + (size_t)(kn + 3) * ldw));
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const float4 a = *reinterpret_cast<const float4 *>(s_h + (row0 + i) * HS + k);
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
        w0 = n0; w1 = n1; w2 = n2; w3 = n3;
    }
}

// Evaluates f_theta (un-normalised head outputs, 9 per row) for the RT rows whose inputs sit in
// sm.x; result lands in sm.out[0][r*12 + c] (c < 9).  `w1col[9]`/`b1v` are this thread's column of
// the first pose-encoder layer (kept in registers by the caller), `s_tq` this stage's t-branch.
// Ends with a __syncthreads(); the caller may read sm.out[0] right after.
template <int RPT>
__device__ __forceinline__ void tile_forward(const float *__restrict__ P, const float *__restrict__ proj,
                                             TileSmem<RPT> &sm, const float *s_tq,
                                             const float (&w1col)[9], float b1v) {
    constexpr int RT = 4 * RPT;
    const int tid = threadIdx.x;
    const int tx = tid & 63, ty = tid >> 6;
    const int row0 = ty * RPT, col = 4 * tx;

    // layer 1: thread n = tid computes column n for every row
    for (int r = 0; r < RT; ++r) {
        const float4 xa = *reinterpret_cast<const float4 *>(sm.x + r * 12);
        const float4 xb = *reinterpret_cast<const float4 *>(sm.x + r * 12 + 4);
        const float xc = sm.x[r * 12 + 8];
        float a = b1v;
        a = fmaf(xa.x, w1col[0], a); a = fmaf(xa.y, w1col[1], a); a = fmaf(xa.z, w1col[2], a);
        a = fmaf(xa.w, w1col[3], a); a = fmaf(xb.x, w1col[4], a); a = fmaf(xb.y, w1col[5], a);
        a = fmaf(xb.z, w1col[6], a); a = fmaf(xb.w, w1col[7], a); a = fmaf(xc, w1col[8], a);
        sm.h1[r * HS + tid] = fmaxf(a, 0.f);
    }
    __syncthreads();

    // layer 2: h2 = relu(h1 @ W2^T + b2)
    {
        float acc[RPT][4];
        const float4 bb = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::B2 + col));
#pragma unroll
        for (int i = 0; i < RPT; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
        gemm256<RPT>(P + TrunkLayout::W2T + col, 256, sm.h1, row0, acc);
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            float4 v = make_float4(fmaxf(acc[i][0], 0.f), fmaxf(acc[i][1], 0.f), fmaxf(acc[i][2], 0.f), fmaxf(acc[i][3], 0.f));
            *reinterpret_cast<float4 *>(sm.h2 + (row0 + i) * HS + col) = v;
        }
    }
    __syncthreads();

    // heads: z = relu(h2 @ Whp^T + proj[obj] + tq);  out = z @ Wo^T (+ bo added by the caller side)
    float part[RPT][9];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int c = 0; c < 9; ++c) part[i][c] = 0.f;
#pragma unroll
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
        gemm256<RPT>(P + TrunkLayout::WHP + h * 256 + col, 768, sm.h2, row0, acc);
        // output layer of this head: 3 dot products over this thread's 4 columns
        const float4 wo0 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 0) * 4));
        const float4 wo1 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 1) * 4));
        const float4 wo2 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 2) * 4));
        const float4 wo3 = __ldg(reinterpret_cast<const float4 *>(P + TrunkLayout::WO + (size_t)(h * 256 + col + 3) * 4));
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
            const float z0 = fmaxf(acc[i][0], 0.f), z1 = fmaxf(acc[i][1], 0.f);
            const float z2 = fmaxf(acc[i][2], 0.f), z3 = fmaxf(acc[i][3], 0.f);
            part[i][h * 3 + 0] = fmaf(z3, wo3.x, fmaf(z2, wo2.x, fmaf(z1, wo1.x, z0 * wo0.x)));
            part[i][h * 3 + 1] = fmaf(z3, wo3.y, fmaf(z2, wo2.y, fmaf(z1, wo1.y, z0 * wo0.y)));
            part[i][h * 3 + 2] = fmaf(z3, wo3.z, fmaf(z2, wo2.z, fmaf(z1, wo1.z, z0 * wo0.z)));
        }
    }
    // reduce the 64 column-threads of each row group: 32 lanes by shuffle, 2 warps through smem
    const int lane = tid & 31, half = (tid >> 5) & 1;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            float v = part[i][c];
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (lane == 0) sm.out[half][(row0 + i) * 12 + c] = v;
        }
    __syncthreads();
    for (int i = tid; i < RT * 9; i += blockDim.x) {
        const int r = i / 9, c = i - 9 * r;
        sm.out[0][r * 12 + c] = (sm.out[0][r * 12 + c] + sm.out[1][r * 12 + c]) + __ldg(P + TrunkLayout::BO + c);
    }
    __syncthreads();
}

// loads this thread's column of W1 (layer 1) into registers
__device__ __forceinline__ void load_w1col(const float *__restrict__ P, float (&w1col)[9], float &b1v) {
    const int n = threadIdx.x;
#pragma unroll
    for (int k = 0; k < 9; ++k) w1col[k] = __ldg(P + TrunkLayout::W1T + k * 256 + n);
    b1v = __ldg(P + TrunkLayout::B1 + n);
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
