// PointNet++ set-abstraction integer ops for sm_100a: farthest-point sampling, ball query,
// gather and grouping.  Indices are bit-exact with the reference extension
// (networks/pts_encoder/pointnet2_utils/pointnet2/src/{sampling,ball_query,group_points}_gpu.cu).
//
// Design (not a port): the reference keeps FPS distances in a global `temp` array, re-reads
// xyz + temp from L1/L2 every iteration and reduces through a log2(BS)-deep shared-memory tree
// with a barrier per level.  Here every point's coordinates and running distance live in
// registers, the cloud is staged once into shared memory with vectorised loads, and the argmax
// is two redux.sync instructions per warp plus ONE block barrier per sampled point.  The
// reference's tie-break (ties are common: clouds are tiled with duplicates) is a property of its
// reduction tree, so it is reproduced as an explicit total order instead of by mimicking the tree.
#include "common.cuh"

namespace gp {

// ------------------------------------------------------------------------------------------
// FPS
// ------------------------------------------------------------------------------------------
// Reference semantics (sampling_gpu.cu:93-209): thread `tid` of BS scans k = tid, tid+BS, ...
// keeping the FIRST strict maximum; the tree then prefers the lower slot at every halving level.
// The overall winner among equal distances is therefore the point with the smallest
//      ord(k) = bitrev_{log2 BS}(k mod BS) * 2^SH + (k div BS)
// (SURVEY.md Appendix A).  We maximise the 64-bit key (float_bits(dist), ~ord).
struct FpsOrder {
    int logBS;   // log2 of the reference block size
    int SH;      // bits reserved for k div BS
    __device__ __forceinline__ unsigned inv_ord(int k) const {
        unsigned low = (unsigned)k & ((1u << logBS) - 1u);
        unsigned rev = logBS ? (__brev(low) >> (32 - logBS)) : 0u;
        unsigned ord = (rev << SH) | ((unsigned)k >> logBS);
        return 0xFFFFFFFFu - ord;
    }
    __device__ __forceinline__ int index_of(unsigned inv) const {
        unsigned ord = 0xFFFFFFFFu - inv;
        unsigned rev = ord >> SH;
        unsigned q = ord & ((1u << SH) - 1u);
        unsigned low = logBS ? (__brev(rev) >> (32 - logBS)) : 0u;
        return (int)((q << logBS) | low);
    }
};

// CHAIN (gp_fps_chain): the encoder samples level k+1 from the centres of level k, i.e. from a cloud that is
// already in FPS order.  FPS of an FPS-ordered cloud is its own prefix 0, 1, ..., m-1 -- the point that maximised
// the running distance over the superset also maximises it over the subset -- unless two candidates tied exactly
// at some step (then the winner depends on the tie-break order, which differs between the levels).  So every run
// records the first step at which the maximum was shared by two different locations (`tie_out`), and a run whose
// input was tie-free for at least m steps (`tie_in[b] >= m`) writes the prefix and returns; anything else samples
// for real.
template <int NWARPS, int PPT, bool COORDS_IN_REGS, bool CHAIN>
__global__ void __launch_bounds__(NWARPS * 32)
fps_kernel(const float *__restrict__ xyz, int N, int m, FpsOrder order, int *__restrict__ idx,
           float *__restrict__ new_xyz, const int *__restrict__ tie_in, int *__restrict__ tie_out, int sel_smem) {
    constexpr int T = NWARPS * 32;
    extern __shared__ __align__(16) float s_raw[];  // [3*N] AoS copy of this object's cloud (+ [m] selected indices)
    // sel_smem: the winners are parked in shared memory (one store per step) and idx / new_xyz are written by all
    // threads after the loop -- thread 0's three dependent loads + four global stores per step are off the critical path
    int *s_sel = reinterpret_cast<int *>(s_raw + ((3 * N + 3) & ~3));
    __shared__ uint2 s_part[2][32];
    __shared__ int s_tie;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float *cloud = xyz + (size_t)b * N * 3;
    int *out = idx + (size_t)b * m;
    float *out_xyz = new_xyz ? new_xyz + (size_t)b * m * 3 : nullptr;

    if (CHAIN) {
        const int free_steps = tie_in ? __ldg(tie_in + b) : 0;
        if (free_steps >= m && m <= N) {
            for (int i = tid; i < m; i += T) out[i] = i;
            if (out_xyz)
                for (int i = tid; i < 3 * m; i += T) out_xyz[i] = __ldg(cloud + i);
            if (tid == 0 && tie_out) tie_out[b] = free_steps;
            return;
        }
        if (tid == 0) s_tie = m;
    }

    // stage the cloud: 16-byte loads when the object's base is aligned (N % 4 == 0)
    const int nfl = N * 3;
    if ((((size_t)b * nfl) & 3) == 0 && (nfl & 3) == 0 && ((uintptr_t)xyz & 15) == 0) {
        const float4 *src = reinterpret_cast<const float4 *>(cloud);
        float4 *dst = reinterpret_cast<float4 *>(s_raw);
        for (int i = tid; i < nfl / 4; i += T) dst[i] = __ldg(src + i);
    } else {
        for (int i = tid; i < nfl; i += T) s_raw[i] = __ldg(cloud + i);
    }
    __syncthreads();

    float px[COORDS_IN_REGS ? PPT : 1], py[COORDS_IN_REGS ? PPT : 1], pz[COORDS_IN_REGS ? PPT : 1];
    float dist[PPT];
    unsigned inv[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int k = tid + i * T;
        const bool valid = k < N;
        if (COORDS_IN_REGS) {
            px[i] = valid ? s_raw[3 * k + 0] : 0.f;
            py[i] = valid ? s_raw[3 * k + 1] : 0.f;
            pz[i] = valid ? s_raw[3 * k + 2] : 0.f;
        }
        dist[i] = valid ? 1e10f : 0.f;  // pointnet2_utils.py:32-34 fills temp with 1e10
        inv[i] = valid ? order.inv_ord(k) : 0u;  // padded lanes can never win: key (0, 0)
    }

    int old = 0;
    unsigned prev_whi = 0u;
    int first_bad = m;
    if (tid == 0) {
        if (sel_smem) {
            s_sel[0] = 0;
        } else {
            out[0] = 0;
            if (out_xyz) {
                out_xyz[0] = s_raw[0];
                out_xyz[1] = s_raw[1];
                out_xyz[2] = s_raw[2];
            }
        }
    }
    const int lane = tid & 31, warp = tid >> 5;
    for (int j = 1; j < m; ++j) {
        const float x1 = s_raw[3 * old + 0], y1 = s_raw[3 * old + 1], z1 = s_raw[3 * old + 2];
        // CHAIN: was the maximum of step j-1 (`dist` still holds that step's distances) held by one location only?
        // Copies of the winner may share it: their distance drops to 0 now, so they are never sampled before the
        // cloud is exhausted.  A holder that stays above 0, or an exhausted cloud (maximum 0), ends the tie-free
        // prefix.
        bool bad = false;
        unsigned best_hi = 0u, best_lo = 0u;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            float x2, y2, z2;
            if (COORDS_IN_REGS) {
                x2 = px[i]; y2 = py[i]; z2 = pz[i];
            } else {
                const int k = min(tid + i * T, N - 1);
                x2 = s_raw[3 * k + 0]; y2 = s_raw[3 * k + 1]; z2 = s_raw[3 * k + 2];
            }
            const float d = sqdist_ref(x2 - x1, y2 - y1, z2 - z1);
            const float d2 = fminf(d, dist[i]);
            if (CHAIN) bad |= __float_as_uint(dist[i]) == prev_whi && d2 != 0.f;
            dist[i] = d2;
            const unsigned hi = __float_as_uint(d2);
            const bool gt = (hi > best_hi) || (hi == best_hi && inv[i] > best_lo);
            best_hi = gt ? hi : best_hi;
            best_lo = gt ? inv[i] : best_lo;
        }
        unsigned whi = __reduce_max_sync(0xffffffffu, best_hi);
        unsigned wlo = __reduce_max_sync(0xffffffffu, best_hi == whi ? best_lo : 0u);
        if (NWARPS > 1) {
            if (lane == 0) s_part[j & 1][warp] = make_uint2(whi, wlo);
            __syncthreads();
            uint2 p = lane < NWARPS ? s_part[j & 1][lane] : make_uint2(0u, 0u);
            whi = __reduce_max_sync(0xffffffffu, p.x);
            wlo = __reduce_max_sync(0xffffffffu, p.x == whi ? p.y : 0u);
        }
        old = order.index_of(wlo);
        if (CHAIN) {  // no vote, no branch: every thread keeps the first bad step it saw
            if (j > 1 && (bad || prev_whi == 0u)) first_bad = min(first_bad, j - 1);
            prev_whi = whi;
        }
        if (tid == 0) {
            if (sel_smem) {
                s_sel[j] = old;
            } else {
                out[j] = old;
                if (out_xyz) {
                    out_xyz[3 * j + 0] = s_raw[3 * old + 0];
                    out_xyz[3 * j + 1] = s_raw[3 * old + 1];
                    out_xyz[3 * j + 2] = s_raw[3 * old + 2];
                }
            }
        }
    }
    if (sel_smem) {
        __syncthreads();
        for (int i = tid; i < m; i += T) out[i] = s_sel[i];
        if (out_xyz)
            for (int i = tid; i < 3 * m; i += T) out_xyz[i] = s_raw[3 * s_sel[i / 3] + i % 3];
    }
    if (CHAIN && tie_out) {  // the last step is not checked: the prefix is vouched for up to m - 1 samples
        first_bad = __reduce_min_sync(0xffffffffu, first_bad);
        if (lane == 0) atomicMin(&s_tie, first_bad);
        __syncthreads();
        if (tid == 0) tie_out[b] = min(s_tie, m - 1);
    }
}

static int ref_block_size(int n) {  // cuda_utils.h:10-14 opt_n_threads
    int p = 0;
    while ((2 << p) <= n) ++p;  // largest power of two <= n
    int bs = 1 << p;
    return bs > 1024 ? 1024 : bs;
}

template <int NWARPS, int PPT, bool REGS>
static int launch_fps(const float *xyz, int B, int N, int m, FpsOrder order, int *idx,
                      float *new_xyz, bool chain, const int *tie_in, int *tie_out, cudaStream_t st) {
    size_t smem = (size_t)N * 3 * sizeof(float);
    smem = (smem + 15) & ~(size_t)15;
    // the selected indices go through shared memory when they fit next to the cloud
    const int sel_smem = smem + (size_t)m * sizeof(int) <= 200 * 1024 ? 1 : 0;
    if (sel_smem) smem += ((size_t)m * sizeof(int) + 15) & ~(size_t)15;
    auto kern = chain ? fps_kernel<NWARPS, PPT, REGS, true> : fps_kernel<NWARPS, PPT, REGS, false>;
    if (smem > 40 * 1024) GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, NWARPS * 32, smem, st>>>(xyz, N, m, order, idx, new_xyz, tie_in, tie_out, sel_smem);
    GP_CHECK_LAUNCH("gp_fps");
    return GP_OK;
}

}  // namespace gp

using namespace gp;

static int fps_dispatch(const char *who, const float *xyz, int B, int N, int m, int32_t *idx, float *new_xyz,
                        bool chain, const int *tie_in, int *tie_out, gp_stream_t s) {
    GP_REQUIRE(B >= 0 && N >= 1 && m >= 0, "%s: bad sizes B=%d N=%d m=%d", who, B, N, m);
    if (B == 0 || m == 0) return GP_OK;
    GP_REQUIRE(xyz && idx, "%s: null pointer", who);
    if (N > 16384) {
        set_error("%s: N=%d > 16384 is not supported (cloud must fit in shared memory)", who, N);
        return GP_ERR_UNSUPPORTED;
    }
    const int BS = ref_block_size(N);
    FpsOrder order;
    order.logBS = 0;
    while ((1 << order.logBS) < BS) ++order.logBS;
    const int J = (N + BS - 1) / BS;
    order.SH = 0;
    while ((1 << order.SH) < J) ++order.SH;
    cudaStream_t st = as_stream(s);
#define GP_FPS_CASE(W, P, R) return launch_fps<W, P, R>(xyz, B, N, m, order, idx, new_xyz, chain, tie_in, tie_out, st)
    if (N <= 32) GP_FPS_CASE(1, 1, true);
    if (N <= 64) GP_FPS_CASE(1, 2, true);
    if (N <= 128) GP_FPS_CASE(1, 4, true);
    if (N <= 256) GP_FPS_CASE(2, 4, true);
    if (N <= 512) GP_FPS_CASE(4, 4, true);
    if (N <= 1024) GP_FPS_CASE(8, 4, true);
    if (N <= 2048) GP_FPS_CASE(16, 4, true);
    if (N <= 4096) GP_FPS_CASE(32, 4, true);
    if (N <= 8192) GP_FPS_CASE(32, 8, false);
    GP_FPS_CASE(32, 16, false);
#undef GP_FPS_CASE
}

extern "C" int gp_fps(const float *xyz, int B, int N, int m, int32_t *idx, float *new_xyz,
                      gp_stream_t s) {
    return fps_dispatch("gp_fps", xyz, B, N, m, idx, new_xyz, false, nullptr, nullptr, s);
}

extern "C" int gp_fps_chain(const float *xyz, int B, int N, int m, int32_t *idx, float *new_xyz,
                            const int32_t *tie_free_in, int32_t *tie_free_out, gp_stream_t s) {
    GP_REQUIRE(tie_free_out || B == 0 || m == 0, "gp_fps_chain: tie_free_out is null");
    return fps_dispatch("gp_fps_chain", xyz, B, N, m, idx, new_xyz, true, tie_free_in, tie_free_out, s);
}

// ------------------------------------------------------------------------------------------
// gather:  out[b,c,j] = points[b,c,idx[b,j]]     (sampling_gpu.cu:8-24)
// ------------------------------------------------------------------------------------------
namespace gp {
__global__ void gather_kernel(const float *__restrict__ points, const int *__restrict__ idx, int C,
                              int N, int m, float *__restrict__ out) {
    const int b = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int id = __ldg(idx + (size_t)b * m + j);
    const float *src = points + (size_t)b * C * N + id;
    float *dst = out + (size_t)b * C * m + j;
    for (int c = 0; c < C; ++c) dst[(size_t)c * m] = __ldg(src + (size_t)c * N);
}
}  // namespace gp

extern "C" int gp_gather(const float *points, const int32_t *idx, int B, int C, int N, int m,
                         float *out, gp_stream_t s) {
    GP_REQUIRE(points && idx && out, "gp_gather: null pointer");
    GP_REQUIRE(B >= 0 && C >= 0 && N >= 1 && m >= 0, "gp_gather: bad sizes");
    if (B == 0 || C == 0 || m == 0) return GP_OK;
    GP_REQUIRE(B <= 65535, "gp_gather: B=%d > 65535", B);
    dim3 grid((m + 127) / 128, B);
    gather_kernel<<<grid, 128, 0, as_stream(s)>>>(points, idx, C, N, m, out);
    GP_CHECK_LAUNCH("gp_gather");
    return GP_OK;
}

// ------------------------------------------------------------------------------------------
// ball query (one or two radii per scan)       (ball_query_gpu.cu:9-45)
// ------------------------------------------------------------------------------------------
namespace gp {

constexpr int BQ_THREADS = 128;                // one centre per thread
constexpr int BQ_CHUNK = 3072;                  // points staged per pass (48 KB as float4)

// One thread per centre, scanning the staged cloud in index order: every lane reads the same float4
// (a shared-memory broadcast), so a point costs one LDS.128 + the reference's three-term distance +
// one compare + one predicated OR per radius into a 32-point hit mask: the scan is branch-free and the
// loads run ahead; the (rare, ~2 % of the pairs) hits are written from the masks once per 32 points.
// The reference's thread-per-centre kernel has the same shape but reads the cloud from global memory and
// branches per point; an earlier warp-per-centre version of this
// kernel (ballot + popc compaction) was latency-bound on its ballot -> count -> branch chain
// (116 us for level 1 at 64 objects, 1 % of HBM bandwidth, profiles/README.md).
// Optional by-products of the scan for the encoder's level buffers (one thread owns one centre): the centre relative
// to a per-object shift, and the [x y z 0] tail row (relative or absolute) the next level's GEMM rows end with.
struct BqTails {
    const float *shift;   // [B,3] or nullptr
    float *rel;           // [B,M,3] = new_xyz - shift, or nullptr
    float *tail;          // [B,M,4] = [rel | 0] (tail_abs == 0) or [new_xyz | 0], or nullptr
    int tail_abs;
};

template <bool TWO>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(const float *__restrict__ new_xyz, const float *__restrict__ xyz, int N, int M,
                  float r0sq, int ns0, int *__restrict__ idx0, float r1sq, int ns1,
                  int *__restrict__ idx1, BqTails tails) {
    extern __shared__ __align__(16) float4 s_pts4[];  // [CH] (x, y, z, -)
    const int CH = min(N, BQ_CHUNK);
    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const float *cloud = xyz + (size_t)b * N * 3;
    const int c = blockIdx.x * BQ_THREADS + tid;
    const bool active = c < M;

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *p = new_xyz + ((size_t)b * M + c) * 3;
        cx = __ldg(p + 0); cy = __ldg(p + 1); cz = __ldg(p + 2);
        if (tails.rel || tails.tail) {
            float sx = 0.f, sy = 0.f, sz = 0.f;
            if (tails.shift) { sx = __ldg(tails.shift + b * 3); sy = __ldg(tails.shift + b * 3 + 1); sz = __ldg(tails.shift + b * 3 + 2); }
            const float rx = cx - sx, ry = cy - sy, rz = cz - sz;
            const size_t bc = (size_t)b * M + c;
            if (tails.rel) { tails.rel[bc * 3] = rx; tails.rel[bc * 3 + 1] = ry; tails.rel[bc * 3 + 2] = rz; }
            if (tails.tail)
                *reinterpret_cast<float4 *>(tails.tail + bc * 4) = tails.tail_abs ? make_float4(cx, cy, cz, 0.f) : make_float4(rx, ry, rz, 0.f);
        }
    }
    int *o0 = idx0 + ((size_t)b * M + (active ? c : 0)) * ns0;
    int *o1 = TWO ? idx1 + ((size_t)b * M + (active ? c : 0)) * ns1 : nullptr;
    int n0 = active ? 0 : ns0, n1 = (TWO && active) ? 0 : ns1;  // idle lanes count as full
    int first0 = 0, first1 = 0;

    for (int base = 0; base < N; base += CH) {
        const int n_here = min(CH, N - base);
        if (base) __syncthreads();
        // stage the chunk: coalesced AoS read -> one float4 per point
        float *sp = reinterpret_cast<float *>(s_pts4);
        for (int i = tid; i < n_here * 3; i += BQ_THREADS) {
            const float v = __ldg(cloud + (size_t)base * 3 + i);
            const int k = i / 3, a = i - 3 * k;
            sp[4 * k + a] = v;
        }
        for (int k = n_here + tid; k < ((n_here + 31) & ~31); k += BQ_THREADS)  // pad: never inside a ball
            s_pts4[k] = make_float4(__int_as_float(0x7f800000), 0.f, 0.f, 0.f);
        __syncthreads();
        for (int k0 = 0; k0 < n_here; k0 += 32) {
            if (__all_sync(0xffffffffu, n0 >= ns0 && (!TWO || n1 >= ns1))) break;  // warp-uniform
            // 32 points, branch-free: one hit bit per point and radius
            unsigned m0 = 0u, m1 = 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 p = s_pts4[k0 + j];
                // reference: (new_x - x)^2 + (new_y - y)^2 + (new_z - z)^2, fma pattern in sqdist_ref
                const float d2 = sqdist_ref(cx - p.x, cy - p.y, cz - p.z);
                m0 |= d2 < r0sq ? 1u << j : 0u;
                if (TWO) m1 |= d2 < r1sq ? 1u << j : 0u;
            }
            // the hits, in index order (about one per lane and group at the encoder's radii)
            while (m0 && n0 < ns0) {
                const int k = base + k0 + __ffs(m0) - 1;
                m0 &= m0 - 1u;
                if (n0 == 0) first0 = k;
                o0[n0++] = k;
            }
            while (TWO && m1 && n1 < ns1) {
                const int k = base + k0 + __ffs(m1) - 1;
                m1 &= m1 - 1u;
                if (n1 == 0) first1 = k;
                o1[n1++] = k;
            }
        }
    }
    if (!active) return;
    // back-fill: first hit replicated into the unused slots; all zeros if the ball is empty
    for (int sl = n0; sl < ns0; ++sl) o0[sl] = first0;
    if (TWO)
        for (int sl = n1; sl < ns1; ++sl) o1[sl] = first1;
}

template <bool TWO>
static int launch_bq(const float *new_xyz, const float *xyz, int B, int N, int M, float r0,
                     int ns0, int *idx0, float r1, int ns1, int *idx1, cudaStream_t st, BqTails tails = BqTails{nullptr, nullptr, nullptr, 0}) {
    const int CH = N < BQ_CHUNK ? N : BQ_CHUNK;
    size_t smem = (size_t)((CH + 31) & ~31) * sizeof(float4);
    dim3 grid((M + BQ_THREADS - 1) / BQ_THREADS, B);
    // radius2 = radius * radius in float32 (ball_query_gpu.cu:23)
    const float r0sq = r0 * r0, r1sq = r1 * r1;
    ball_query_kernel<TWO><<<grid, BQ_THREADS, smem, st>>>(new_xyz, xyz, N, M, r0sq, ns0, idx0,
                                                              r1sq, ns1, idx1, tails);
    GP_CHECK_LAUNCH("gp_ball_query");
    return GP_OK;
}
}  // namespace gp

extern "C" int gp_ball_query(const float *new_xyz, const float *xyz, int B, int N, int M,
                             float radius, int nsample, int32_t *idx, gp_stream_t s) {
    GP_REQUIRE(new_xyz && xyz && idx, "gp_ball_query: null pointer");
    GP_REQUIRE(B >= 0 && N >= 1 && M >= 0 && nsample >= 1, "gp_ball_query: bad sizes");
    if (B == 0 || M == 0) return GP_OK;
    GP_REQUIRE(B <= 65535, "gp_ball_query: B=%d > 65535", B);
    return launch_bq<false>(new_xyz, xyz, B, N, M, radius, nsample, idx, 0.f, 0, nullptr, as_stream(s));
}

extern "C" int gp_ball_query2(const float *new_xyz, const float *xyz, int B, int N, int M,
                              float radius0, int nsample0, int32_t *idx0, float radius1,
                              int nsample1, int32_t *idx1, gp_stream_t s) {
    GP_REQUIRE(new_xyz && xyz && idx0 && idx1, "gp_ball_query2: null pointer");
    GP_REQUIRE(B >= 0 && N >= 1 && M >= 0 && nsample0 >= 1 && nsample1 >= 1, "gp_ball_query2: bad sizes");
    if (B == 0 || M == 0) return GP_OK;
    GP_REQUIRE(B <= 65535, "gp_ball_query2: B=%d > 65535", B);
    return launch_bq<true>(new_xyz, xyz, B, N, M, radius0, nsample0, idx0, radius1, nsample1, idx1,
                           as_stream(s));
}

extern "C" int gp_ball_query2_tails(const float *new_xyz, const float *xyz, int B, int N, int M,
                                    float radius0, int nsample0, int32_t *idx0, float radius1,
                                    int nsample1, int32_t *idx1, const float *shift, float *rel, float *tail,
                                    int tail_absolute, gp_stream_t s) {
    GP_REQUIRE(new_xyz && xyz && idx0 && idx1, "gp_ball_query2_tails: null pointer");
    GP_REQUIRE(B >= 0 && N >= 1 && M >= 0 && nsample0 >= 1 && nsample1 >= 1, "gp_ball_query2_tails: bad sizes");
    GP_REQUIRE(!tail || ((uintptr_t)tail & 15) == 0, "gp_ball_query2_tails: tail must be 16-byte aligned");
    if (B == 0 || M == 0) return GP_OK;
    GP_REQUIRE(B <= 65535, "gp_ball_query2_tails: B=%d > 65535", B);
    return launch_bq<true>(new_xyz, xyz, B, N, M, radius0, nsample0, idx0, radius1, nsample1, idx1,
                           as_stream(s), BqTails{shift, rel, tail, tail_absolute});
}

// ------------------------------------------------------------------------------------------
// grouping: out[b,c,p,s] = points[b,c,idx[b,p,s]]          (group_points_gpu.cu:47-66)
// and the fused QueryAndGroup tail (pointnet2_utils.py:279-296)
// ------------------------------------------------------------------------------------------
namespace gp {

constexpr int GR_CC = 8;  // channels per thread (amortises the idx load)

// One thread owns 4 consecutive (p,s) slots (16-byte store) for GR_CC channels.
// FUSED: channel 0..2 = xyz[idx] - new_xyz, channel 3.. = features[idx].
template <bool FUSED>
__global__ void __launch_bounds__(256)
group_kernel(const float *__restrict__ points, const float *__restrict__ xyz,
             const float *__restrict__ new_xyz, const int *__restrict__ idx, int C, int N, int M,
             int ns, float *__restrict__ out) {
    const int b = blockIdx.z;
    const int total = M * ns;  // multiple of 4 (checked on the host)
    const int q4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q4 >= total) return;
    const int4 id = __ldg(reinterpret_cast<const int4 *>(idx + (size_t)b * total + q4));
    const int Ctot = FUSED ? C + 3 : C;
    const int c0 = blockIdx.y * GR_CC;
    const int c1 = min(c0 + GR_CC, Ctot);
    for (int c = c0; c < c1; ++c) {
        float4 v;
        if (FUSED && c < 3) {
            const float *px = xyz + (size_t)b * N * 3 + c;
            // ns % 4 == 0 here, so the four slots share one centre
            const float ctr = __ldg(new_xyz + ((size_t)b * M + q4 / ns) * 3 + c);
            v.x = __ldg(px + 3 * id.x) - ctr;
            v.y = __ldg(px + 3 * id.y) - ctr;
            v.z = __ldg(px + 3 * id.z) - ctr;
            v.w = __ldg(px + 3 * id.w) - ctr;
        } else {
            const float *src = points + ((size_t)b * C + (FUSED ? c - 3 : c)) * N;
            v.x = __ldg(src + id.x);
            v.y = __ldg(src + id.y);
            v.z = __ldg(src + id.z);
            v.w = __ldg(src + id.w);
        }
        *reinterpret_cast<float4 *>(out + ((size_t)b * Ctot + c) * total + q4) = v;
    }
}

// scalar fallback for M*ns (or ns, fused) not divisible by 4
template <bool FUSED>
__global__ void group_kernel_scalar(const float *__restrict__ points, const float *__restrict__ xyz,
                                    const float *__restrict__ new_xyz, const int *__restrict__ idx,
                                    int C, int N, int M, int ns, float *__restrict__ out) {
    const int b = blockIdx.z;
    const int total = M * ns;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    const int id = __ldg(idx + (size_t)b * total + q);
    const int Ctot = FUSED ? C + 3 : C;
    const int c0 = blockIdx.y * GR_CC, c1 = min(c0 + GR_CC, Ctot);
    for (int c = c0; c < c1; ++c) {
        float v;
        if (FUSED && c < 3)
            v = __ldg(xyz + ((size_t)b * N + id) * 3 + c) - __ldg(new_xyz + ((size_t)b * M + q / ns) * 3 + c);
        else
            v = __ldg(points + ((size_t)b * C + (FUSED ? c - 3 : c)) * N + id);
        out[((size_t)b * Ctot + c) * total + q] = v;
    }
}

template <bool FUSED>
static int launch_group(const float *points, const float *xyz, const float *new_xyz, const int *idx,
                        int B, int C, int N, int M, int ns, float *out, cudaStream_t st) {
    const int Ctot = FUSED ? C + 3 : C;
    const int total = M * ns;
    const bool vec = (ns % 4 == 0) && (((uintptr_t)idx & 15) == 0) && (((uintptr_t)out & 15) == 0);
    if (vec) {
        dim3 grid((total / 4 + 255) / 256, (Ctot + GR_CC - 1) / GR_CC, B);
        group_kernel<FUSED><<<grid, 256, 0, st>>>(points, xyz, new_xyz, idx, C, N, M, ns, out);
    } else {
        dim3 grid((total + 255) / 256, (Ctot + GR_CC - 1) / GR_CC, B);
        group_kernel_scalar<FUSED><<<grid, 256, 0, st>>>(points, xyz, new_xyz, idx, C, N, M, ns, out);
    }
    GP_CHECK_LAUNCH(FUSED ? "gp_query_group" : "gp_group");
    return GP_OK;
}
}  // namespace gp

extern "C" int gp_group(const float *points, const int32_t *idx, int B, int C, int N, int M,
                        int nsample, float *out, gp_stream_t s) {
    GP_REQUIRE(points && idx && out, "gp_group: null pointer");
    GP_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0 && nsample >= 0, "gp_group: bad sizes");
    if (B == 0 || C == 0 || M == 0 || nsample == 0) return GP_OK;
    GP_REQUIRE(B <= 65535 && (C + GR_CC - 1) / GR_CC <= 65535, "gp_group: B or C too large");
    return launch_group<false>(points, nullptr, nullptr, idx, B, C, N, M, nsample, out, as_stream(s));
}

extern "C" int gp_query_group(const float *xyz, const float *new_xyz, const float *features,
                              const int32_t *idx, int B, int C, int N, int M, int nsample,
                              float *out, gp_stream_t s) {
    GP_REQUIRE(xyz && new_xyz && idx && out, "gp_query_group: null pointer");
    GP_REQUIRE(C == 0 || features, "gp_query_group: features is NULL but C=%d", C);
    GP_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0 && nsample >= 0, "gp_query_group: bad sizes");
    if (B == 0 || M == 0 || nsample == 0) return GP_OK;
    GP_REQUIRE(B <= 65535, "gp_query_group: B too large");
    return launch_group<true>(features, xyz, new_xyz, idx, B, C, N, M, nsample, out, as_stream(s));
}

// ------------------------------------------------------------------------------------------
// channels-last ("rows") variants used by the drop-in encoder: one GEMM row per (centre, sample)
// ------------------------------------------------------------------------------------------
namespace gp {

// out[(b*M + p)*ns + s][0:3] = xyz[b, idx[b,p,s]] - new_xyz[b,p];  out[..][3:3+C] = feat[b, idx[b,p,s], :]
// (the same values QueryAndGroup.forward produces, P2/pointnet2_utils.py:279-296, stored row-major per
// sample instead of [B, 3+C, M, ns]); columns 3+C .. ld-1 are zero-filled.
// One warp per row: lanes copy the feature row with 16-byte accesses when C % 4 == 0.
__global__ void __launch_bounds__(256)
group_rows_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                  const float *__restrict__ feat, const int *__restrict__ idx, int C, int N, int M, int ns,
                  int ld, long long rows_total, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows_total) return;
    const long long bp = row / ns;           // b*M + p
    const int b = (int)(bp / M);
    const int id = __ldg(idx + row);
    float *o = out + row * ld;
    if (lane < 3) o[lane] = __ldg(xyz + ((size_t)b * N + id) * 3 + lane) - __ldg(new_xyz + bp * 3 + lane);
    if (C > 0) {
        const float *src = feat + ((size_t)b * N + id) * C;
        // destination starts at column 3: scalar stores (rows are short), 16-byte loads when aligned
        if ((C & 3) == 0 && (((uintptr_t)feat) & 15) == 0) {
            for (int c4 = lane; c4 < C / 4; c4 += 32) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(src) + c4);
                float *d = o + 3 + 4 * c4;
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
        } else {
            for (int c = lane; c < C; c += 32) o[3 + c] = __ldg(src + c);
        }
    }
    for (int c = 3 + C + lane; c < ld; c += 32) o[c] = 0.f;
}

// out[g*ld_out + c] = max_s h[(g*ns + s)*C + c]      (F.max_pool2d over nsample, pointnet2_modules.py:59-61)
__global__ void __launch_bounds__(256)
maxpool_rows_kernel(const float *__restrict__ h, long long G, int ns, int C, int ld_out, float *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over G * C/4 (vector) or G*C
    const int C4 = C >> 2;
    if ((C & 3) == 0 && (ld_out & 3) == 0 && ((uintptr_t)h & 15) == 0 && ((uintptr_t)out & 15) == 0) {
        if (i >= G * C4) return;
        const long long g = i / C4;
        const int c4 = (int)(i - g * C4);
        const float4 *p = reinterpret_cast<const float4 *>(h + g * ns * (long long)C) + c4;
        float4 m = __ldg(p);
        for (int s = 1; s < ns; ++s) {
            const float4 v = __ldg(p + (size_t)s * C4);
            m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
        }
        *reinterpret_cast<float4 *>(out + g * ld_out + 4 * c4) = m;
    } else {
        if (i >= G * C) return;
        const long long g = i / C;
        const int c = (int)(i - g * C);
        float m = h[g * ns * (long long)C + c];
        for (int s = 1; s < ns; ++s) m = fmaxf(m, h[(g * ns + s) * (long long)C + c]);
        out[g * ld_out + c] = m;
    }
}
}  // namespace gp

extern "C" int gp_group_rows(const float *xyz, const float *new_xyz, const float *feat_cl, const int32_t *idx,
                             int B, int C, int N, int M, int nsample, int ld_out, float *out, gp_stream_t s) {
    GP_REQUIRE(B >= 0 && C >= 0 && N >= 1 && M >= 0 && nsample >= 0 && ld_out >= 3 + C, "gp_group_rows: bad sizes");
    const long long rows = (long long)B * M * nsample;
    if (rows == 0) return GP_OK;
    GP_REQUIRE(xyz && new_xyz && idx && out && (C == 0 || feat_cl), "gp_group_rows: null pointer");
    GP_REQUIRE((rows + 7) / 8 < 2147483647LL, "gp_group_rows: too many rows");
    group_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(s)>>>(xyz, new_xyz, feat_cl, idx, C, N, M, nsample,
                                                                            ld_out, rows, out);
    GP_CHECK_LAUNCH("gp_group_rows");
    return GP_OK;
}

extern "C" int gp_maxpool_rows(const float *h, long long G, int nsample, int C, int ld_out, float *out, gp_stream_t s) {
    GP_REQUIRE(G >= 0 && nsample >= 1 && C >= 1 && ld_out >= C, "gp_maxpool_rows: bad sizes");
    if (G == 0) return GP_OK;
    GP_REQUIRE(h && out, "gp_maxpool_rows: null pointer");
    const long long work = ((C & 3) == 0 && (ld_out & 3) == 0) ? G * (C >> 2) : G * C;
    maxpool_rows_kernel<<<(unsigned)((work + 255) / 256), 256, 0, as_stream(s)>>>(h, G, nsample, C, ld_out, out);
    GP_CHECK_LAUNCH("gp_maxpool_rows");
    return GP_OK;
}

// ------------------------------------------------------------------------------------------
// First set-abstraction level fused end to end (no point features yet: 3 input channels):
// gather (xyz[idx] - new_xyz) -> 3 -> C1 -> C2 -> C3 SharedMLP (BatchNorm folded, ReLU) -> max over
// the nsample rows of each centre.  One thread per (centre, sample) row, FP32 FFMA, weights broadcast
// from shared memory as float4, pooling with one redux.sync per channel.  The channels are far too
// narrow for a tensor-core tile (16..64), and nothing is materialised in HBM.
// ------------------------------------------------------------------------------------------
namespace gp {

template <int C1, int C2, int C3>
struct SaSmallSmem {
    float w0[3][C1], b0[C1];
    float w1[C1][C2], b1[C2];
    float w2[C2][C3], b2[C3];
};

template <int CIN, int COUT>
__device__ __forceinline__ void dense_relu(const float (&in)[CIN], const float *w /*[CIN][COUT]*/, const float *b,
                                           float (&out)[COUT]) {
#pragma unroll
    for (int n = 0; n < COUT; n += 4) {
        const float4 bv = *reinterpret_cast<const float4 *>(b + n);
        out[n] = bv.x; out[n + 1] = bv.y; out[n + 2] = bv.z; out[n + 3] = bv.w;
    }
#pragma unroll
    for (int k = 0; k < CIN; ++k) {
#pragma unroll
        for (int n = 0; n < COUT; n += 4) {
            const float4 wv = *reinterpret_cast<const float4 *>(w + k * COUT + n);
            out[n] = fmaf(in[k], wv.x, out[n]);
            out[n + 1] = fmaf(in[k], wv.y, out[n + 1]);
            out[n + 2] = fmaf(in[k], wv.z, out[n + 2]);
            out[n + 3] = fmaf(in[k], wv.w, out[n + 3]);
        }
    }
#pragma unroll
    for (int n = 0; n < COUT; ++n) out[n] = fmaxf(out[n], 0.f);
}

template <int C1, int C2, int C3, int NS>
__global__ void __launch_bounds__(256)
sa_small_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, const int *__restrict__ idx,
                int N, int M, long long rows_total, const float *__restrict__ w0, const float *__restrict__ b0,
                const float *__restrict__ w1, const float *__restrict__ b1, const float *__restrict__ w2,
                const float *__restrict__ b2, float *__restrict__ out, int ld_out) {
    __shared__ __align__(16) SaSmallSmem<C1, C2, C3> S;
    const int tid = threadIdx.x, lane = tid & 31;
    // stage the (folded) weights transposed to [k][n]
    for (int i = tid; i < 3 * C1; i += 256) S.w0[i / C1][i % C1] = __ldg(w0 + (i % C1) * 3 + i / C1);
    for (int i = tid; i < C1 * C2; i += 256) S.w1[i / C2][i % C2] = __ldg(w1 + (i % C2) * C1 + i / C2);
    for (int i = tid; i < C2 * C3; i += 256) S.w2[i / C3][i % C3] = __ldg(w2 + (i % C3) * C2 + i / C3);
    for (int i = tid; i < C1; i += 256) S.b0[i] = __ldg(b0 + i);
    for (int i = tid; i < C2; i += 256) S.b1[i] = __ldg(b1 + i);
    for (int i = tid; i < C3; i += 256) S.b2[i] = __ldg(b2 + i);
    __syncthreads();
    const long long row = (long long)blockIdx.x * 256 + tid;
    const bool ok = row < rows_total;
    const long long rr = ok ? row : rows_total - 1;
    const long long bp = rr / NS;           // b*M + p
    const int b = (int)(bp / M);
    const int id = __ldg(idx + rr);
    float a0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) a0[c] = __ldg(xyz + ((size_t)b * N + id) * 3 + c) - __ldg(new_xyz + bp * 3 + c);
    float h1[C1], h2[C2], h3[C3];
    dense_relu<3, C1>(a0, &S.w0[0][0], S.b0, h1);
    dense_relu<C1, C2>(h1, &S.w1[0][0], S.b1, h2);
    dense_relu<C2, C3>(h2, &S.w2[0][0], S.b2, h3);
    // max over the NS rows of the centre: values >= 0, so unsigned order of the bit patterns == float order
    constexpr int GL = NS < 32 ? NS : 32;
    const unsigned mask = GL == 32 ? 0xffffffffu : (((1u << GL) - 1u) << ((lane / GL) * GL));
    float keep[C3 / GL > 0 ? C3 / GL : 1];
#pragma unroll
    for (int n = 0; n < C3; ++n) {
        const unsigned m = redux_max_u32(mask, ok ? __float_as_uint(h3[n]) : 0u);
        if ((lane % GL) == (n % GL)) keep[n / GL] = __uint_as_float(m);
    }
    if (ok) {
        float *dst = out + bp * (long long)ld_out;
#pragma unroll
        for (int q = 0; q < C3 / GL; ++q) dst[q * GL + (lane % GL)] = keep[q];
    }
}

// The same scale with the weights in the kernel's parameter space (constant bank): every weight is the same for
// all lanes, so FFMA takes it as a constant operand and the LDS.128 per four FFMAs disappears.  The host passes
// HOST copies of the folded weights (they are copied into the launch, private to it -- no staging kernel, no race
// between the two encoders' launches).
template <int C1, int C2, int C3>
struct SaSmallConst {
    float w0[3 * C1], b0[C1];      // [k][n]
    float w1[C1 * C2], b1[C2];
    float w2[C2 * C3], b2[C3];
};

template <int CIN, int COUT>
__device__ __forceinline__ void dense_relu_c(const float (&in)[CIN], const float *w /*[CIN][COUT], constant*/,
                                             const float *b, float (&out)[COUT]) {
#pragma unroll
    for (int n = 0; n < COUT; ++n) out[n] = b[n];
#pragma unroll
    for (int k = 0; k < CIN; ++k)
#pragma unroll
        for (int n = 0; n < COUT; ++n) out[n] = fmaf(in[k], w[k * COUT + n], out[n]);
#pragma unroll
    for (int n = 0; n < COUT; ++n) out[n] = fmaxf(out[n], 0.f);
}

template <int C1, int C2, int C3, int NS>
__global__ void __launch_bounds__(256)
sa_small_const_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, const int *__restrict__ idx,
                      int N, int M, long long rows_total, const __grid_constant__ SaSmallConst<C1, C2, C3> W,
                      float *__restrict__ out, int ld_out, const float *__restrict__ tail_src, int tail_col) {
    const int tid = threadIdx.x, lane = tid & 31;
    const long long row = (long long)blockIdx.x * 256 + tid;
    const bool ok = row < rows_total;
    const long long rr = ok ? row : rows_total - 1;
    const long long bp = rr / NS;           // b*M + p
    const int b = (int)(bp / M);
    const int id = __ldg(idx + rr);
    float a0[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) a0[c] = __ldg(xyz + ((size_t)b * N + id) * 3 + c) - __ldg(new_xyz + bp * 3 + c);
    float h1[C1], h2[C2], h3[C3];
    dense_relu_c<3, C1>(a0, W.w0, W.b0, h1);
    dense_relu_c<C1, C2>(h1, W.w1, W.b1, h2);
    dense_relu_c<C2, C3>(h2, W.w2, W.b2, h3);
    constexpr int GL = NS < 32 ? NS : 32;
    const unsigned mask = GL == 32 ? 0xffffffffu : (((1u << GL) - 1u) << ((lane / GL) * GL));
    float keep[C3 / GL > 0 ? C3 / GL : 1];
#pragma unroll
    for (int n = 0; n < C3; ++n) {
        const unsigned m = redux_max_u32(mask, ok ? __float_as_uint(h3[n]) : 0u);
        if ((lane % GL) == (n % GL)) keep[n / GL] = __uint_as_float(m);
    }
    if (ok) {
        float *dst = out + bp * (long long)ld_out;
#pragma unroll
        for (int q = 0; q < C3 / GL; ++q) dst[q * GL + (lane % GL)] = keep[q];
        // the centre's [x y z 0] row for the tail of the level buffer (what the next level's GEMM rows end with)
        if (tail_src && (lane % GL) < 4) dst[tail_col + (lane % GL)] = __ldg(tail_src + bp * 4 + (lane % GL));
    }
}

template <int C1, int C2, int C3>
static int launch_sa_small_const(const float *xyz, const float *new_xyz, const int *idx, int B, int N, int M, int ns,
                                 const float *const *w, const float *const *b, float *out, int ld_out, cudaStream_t st,
                                 const float *tail_src = nullptr, int tail_col = 0) {
    SaSmallConst<C1, C2, C3> W;   // host weights [Cout][Cin] -> [k][n]
    for (int n = 0; n < C1; ++n) { for (int k = 0; k < 3; ++k) W.w0[k * C1 + n] = w[0][n * 3 + k]; W.b0[n] = b[0][n]; }
    for (int n = 0; n < C2; ++n) { for (int k = 0; k < C1; ++k) W.w1[k * C2 + n] = w[1][n * C1 + k]; W.b1[n] = b[1][n]; }
    for (int n = 0; n < C3; ++n) { for (int k = 0; k < C2; ++k) W.w2[k * C3 + n] = w[2][n * C2 + k]; W.b2[n] = b[2][n]; }
    const long long rows = (long long)B * M * ns;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    if (ns == 16)
        sa_small_const_kernel<C1, C2, C3, 16><<<grid, 256, 0, st>>>(xyz, new_xyz, idx, N, M, rows, W, out, ld_out, tail_src, tail_col);
    else
        sa_small_const_kernel<C1, C2, C3, 32><<<grid, 256, 0, st>>>(xyz, new_xyz, idx, N, M, rows, W, out, ld_out, tail_src, tail_col);
    GP_CHECK_LAUNCH("gp_sa_small_mlp_hostw");
    return GP_OK;
}

template <int C1, int C2, int C3>
static int launch_sa_small(const float *xyz, const float *new_xyz, const int *idx, int B, int N, int M, int ns,
                           const float *const *w, const float *const *b, float *out, int ld_out, cudaStream_t st) {
    const long long rows = (long long)B * M * ns;
    const unsigned grid = (unsigned)((rows + 255) / 256);
    if (ns == 16)
        sa_small_kernel<C1, C2, C3, 16><<<grid, 256, 0, st>>>(xyz, new_xyz, idx, N, M, rows, w[0], b[0], w[1], b[1], w[2], b[2], out, ld_out);
    else
        sa_small_kernel<C1, C2, C3, 32><<<grid, 256, 0, st>>>(xyz, new_xyz, idx, N, M, rows, w[0], b[0], w[1], b[1], w[2], b[2], out, ld_out);
    GP_CHECK_LAUNCH("gp_sa_small_mlp");
    return GP_OK;
}
}  // namespace gp

extern "C" int gp_sa_small_mlp(const float *xyz, const float *new_xyz, const int32_t *idx, int B, int N, int M,
                               int nsample, const float *const *weights, const float *const *biases, int C1, int C2,
                               int C3, float *out, int ld_out, gp_stream_t s) {
    GP_REQUIRE(B >= 0 && N >= 1 && M >= 0 && (nsample == 16 || nsample == 32), "gp_sa_small_mlp: nsample must be 16 or 32");
    if ((long long)B * M == 0) return GP_OK;
    GP_REQUIRE(xyz && new_xyz && idx && weights && biases && out && ld_out >= C3, "gp_sa_small_mlp: null pointer / bad ld_out");
    for (int i = 0; i < 3; ++i) GP_REQUIRE(weights[i] && biases[i], "gp_sa_small_mlp: null layer %d", i);
    if (C1 == 16 && C2 == 16 && C3 == 32)
        return launch_sa_small<16, 16, 32>(xyz, new_xyz, idx, B, N, M, nsample, weights, biases, out, ld_out, as_stream(s));
    if (C1 == 32 && C2 == 32 && C3 == 64)
        return launch_sa_small<32, 32, 64>(xyz, new_xyz, idx, B, N, M, nsample, weights, biases, out, ld_out, as_stream(s));
    set_error("gp_sa_small_mlp: channel spec %d-%d-%d is not instantiated (16-16-32, 32-32-64)", C1, C2, C3);
    return GP_ERR_UNSUPPORTED;
}

extern "C" int gp_sa_small_mlp_hostw(const float *xyz, const float *new_xyz, const int32_t *idx, int B, int N, int M,
                                     int nsample, const float *const *host_weights, const float *const *host_biases,
                                     int C1, int C2, int C3, float *out, int ld_out, gp_stream_t s) {
    GP_REQUIRE(B >= 0 && N >= 1 && M >= 0 && (nsample == 16 || nsample == 32), "gp_sa_small_mlp_hostw: nsample must be 16 or 32");
    if ((long long)B * M == 0) return GP_OK;
    GP_REQUIRE(xyz && new_xyz && idx && host_weights && host_biases && out && ld_out >= C3,
               "gp_sa_small_mlp_hostw: null pointer / bad ld_out");
    for (int i = 0; i < 3; ++i) GP_REQUIRE(host_weights[i] && host_biases[i], "gp_sa_small_mlp_hostw: null layer %d", i);
    if (C1 == 16 && C2 == 16 && C3 == 32)
        return gp::launch_sa_small_const<16, 16, 32>(xyz, new_xyz, idx, B, N, M, nsample, host_weights, host_biases, out, ld_out, gp::as_stream(s));
    if (C1 == 32 && C2 == 32 && C3 == 64)
        return gp::launch_sa_small_const<32, 32, 64>(xyz, new_xyz, idx, B, N, M, nsample, host_weights, host_biases, out, ld_out, gp::as_stream(s));
    gp::set_error("gp_sa_small_mlp_hostw: channel spec %d-%d-%d is not instantiated (16-16-32, 32-32-64)", C1, C2, C3);
    return GP_ERR_UNSUPPORTED;
}

extern "C" int gp_sa_small_mlp_hostw_tail(const float *xyz, const float *new_xyz, const int32_t *idx, int B, int N, int M,
                                          int nsample, const float *const *host_weights, const float *const *host_biases,
                                          int C1, int C2, int C3, float *out, int ld_out, const float *tail_src,
                                          int tail_col, gp_stream_t s) {
    GP_REQUIRE(B >= 0 && N >= 1 && M >= 0 && (nsample == 16 || nsample == 32), "gp_sa_small_mlp_hostw_tail: nsample must be 16 or 32");
    if ((long long)B * M == 0) return GP_OK;
    GP_REQUIRE(xyz && new_xyz && idx && host_weights && host_biases && out && ld_out >= C3,
               "gp_sa_small_mlp_hostw_tail: null pointer / bad ld_out");
    GP_REQUIRE(!tail_src || (tail_col >= 0 && tail_col + 4 <= ld_out), "gp_sa_small_mlp_hostw_tail: the tail does not fit the row");
    for (int i = 0; i < 3; ++i) GP_REQUIRE(host_weights[i] && host_biases[i], "gp_sa_small_mlp_hostw_tail: null layer %d", i);
    if (C1 == 16 && C2 == 16 && C3 == 32)
        return gp::launch_sa_small_const<16, 16, 32>(xyz, new_xyz, idx, B, N, M, nsample, host_weights, host_biases, out, ld_out, gp::as_stream(s), tail_src, tail_col);
    if (C1 == 32 && C2 == 32 && C3 == 64)
        return gp::launch_sa_small_const<32, 32, 64>(xyz, new_xyz, idx, B, N, M, nsample, host_weights, host_biases, out, ld_out, gp::as_stream(s), tail_src, tail_col);
    gp::set_error("gp_sa_small_mlp_hostw_tail: channel spec %d-%d-%d is not instantiated (16-16-32, 32-32-64)", C1, C2, C3);
    return GP_ERR_UNSUPPORTED;
}
