// Layers 2 and 3 of a PointNet++ set-abstraction scale and its max-pool in ONE persistent tensor-core kernel
// (P2/pointnet2_modules.py:45-66 on the rows produced by P2/pointnet2_utils.py:279-296):
//
//   A[r][k] = relu(P[batch(r) * n_src + gidx[r]][k] - Q[r / q_ns][k])          hoisted first layer, applied on the fly
//   H       = relu(A . W1^T + b1)                                              [128 rows x c2], never leaves the SM
//   out[g]  = max over the pool_ns rows of group g of relu(H . W2^T + b2)      [rows / pool_ns, c3]
//
// The (centre, sample) activation matrices of the reference -- c2 and c3 floats for each of B*npoint*nsample
// rows -- never exist in HBM: a tile of 128 rows is gathered from the (L2 resident) per-point table P,
// converted to bf16 (hi / lo) A-operand atoms in shared memory, multiplied on tcgen05 into TMEM, re-packed as
// the next A operand by the epilogue warps, multiplied again and pooled straight out of TMEM.
//
// CTA = 10 warps, persistent over tiles: warp 0 streams the weight chunks (cp.async.bulk, [<=128 n][64 k]
// images from the gp_gemm_pack layout) through a shared-memory ring and runs ahead across tiles; warp 1 issues
// the MMAs (M128, N <= 128 per chunk, K16); warps 2..9 gather / convert / run both epilogues.
// NPASS = 1: bf16 operands; NPASS = 3: split-bf16 (hi*hi + lo*hi + hi*lo), fp32-class.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace gp {
namespace saf {

using namespace gp::tc;

constexpr int BM = 128;
constexpr int ATOM = 128 * 128;   // [128 rows][64 k] bf16
constexpr int HALF = 128 * 128;   // weight chunk: up to [128 n][64 k] bf16

struct Args {
    const float *P;
    int n_src, ldp;
    const int *gidx;
    long long R;
    int rows_per_batch;
    const float *Q;
    int ldq, q_ns;
    const uint8_t *W1p;
    const float *b1;
    int c1, c2;
    const uint8_t *W2p;
    const float *b2;
    int c3;
    int pool_ns;
    float *pooled;
    int ld_pooled;
    int ntiles;
    // launch configuration (host): ring stages, bytes per weight image slot, TMEM columns, TMEM column of D2
    int nst, slot_bytes, tmem_cols, d2col;
    int chunk_rows;      // output columns per weight chunk: 128, or 64 when shared memory is tight
    int bias_smem;       // biases staged in shared memory (else read through L1)
    int tile_in_batch;   // rows_per_batch % 128 == 0: all rows of a tile belong to one batch
    int q_shift;         // log2(q_ns) if q_ns is a power of two, else -1
};

template <int NPASS>
struct Cfg {
    static constexpr int IMAGES = NPASS == 3 ? 2 : 1;
};
constexpr int MAX_NST = 4;

__host__ __device__ inline int round16(int n) { return (n + 15) & ~15; }
__host__ __device__ inline int atoms_of(int k) { return (k + 63) / 64; }

// shared memory carve-up (bytes from a 1024-aligned base): weight ring | A buffer (= pool staging) | biases | barriers
template <int NPASS>
struct Layout {
    int natoms, nst, slot, bias_smem;
    __host__ __device__ Layout(int c1, int c2, int nst_, int slot_, int bias_smem_)
        : natoms(atoms_of(c1) > atoms_of(c2) ? atoms_of(c1) : atoms_of(c2)), nst(nst_), slot(slot_), bias_smem(bias_smem_) {}
    __host__ __device__ size_t ring() const { return 0; }
    __host__ __device__ size_t abuf() const { return (size_t)nst * Cfg<NPASS>::IMAGES * slot; }
    // the A buffer doubles as the pooling stage (8 x [32][32] f32) once the second GEMM has read it
    __host__ __device__ size_t abuf_bytes() const {
        const size_t b = (size_t)Cfg<NPASS>::IMAGES * natoms * ATOM;
        return b < 8 * 4096 ? 8 * 4096 : b;
    }
    __host__ __device__ size_t bias() const { return abuf() + abuf_bytes(); }
    __host__ __device__ size_t bars() const { return bias() + (bias_smem ? (384 + 512) * sizeof(float) : 0); }
    __host__ __device__ size_t total() const { return bars() + 256 + 1024; }
};

// Max-pool of one warp's 32 rows x 32 columns block of raw accumulators (v = this lane's row; -inf for rows that
// do not exist) over groups of `ns` consecutive rows, then bias + ReLU once per pooled value: the bias is per
// column and x -> relu(x + b) is monotone, so relu(max_r(x_r) + b) == max_r(relu(x_r + b)) bit for bit.
// Transposed through a swizzled [32][32] shared-memory tile (bank = col ^ row on both sides); lane c reduces
// column c and stores it -- no cross-lane instructions, coalesced 128-byte stores.
__device__ __forceinline__ void pool_block(const float (&v)[32], float *stg, int lane, long long grow0, long long R, int ns,
                                           int nbase, int N, const float *bias, float *pooled, int ld_pooled) {
#pragma unroll
    for (int j = 0; j < 32; ++j) stg[lane * 32 + (j ^ lane)] = v[j];
    __syncwarp();
    const int n = nbase + lane;
    const float bn = n < N ? bias[n] : 0.f;
    float m8[4];  // maxima of rows 0-7, 8-15, 16-23, 24-31 of column `lane` (all loads independent)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = stg[(8 * q + i) * 32 + (lane ^ (8 * q + i))];
        m8[q] = fmaxf(fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3])), fmaxf(fmaxf(t[4], t[5]), fmaxf(t[6], t[7])));
    }
    if (n < N) {
        if (ns == 32) {
            if (grow0 < R) pooled[(grow0 / 32) * (long long)ld_pooled + n] = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])) + bn, 0.f);
        } else if (ns == 16) {
            if (grow0 < R) pooled[(grow0 / 16) * (long long)ld_pooled + n] = fmaxf(fmaxf(m8[0], m8[1]) + bn, 0.f);
            if (grow0 + 16 < R) pooled[(grow0 / 16 + 1) * (long long)ld_pooled + n] = fmaxf(fmaxf(m8[2], m8[3]) + bn, 0.f);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (grow0 + 8 * q < R) pooled[(grow0 / 8 + q) * (long long)ld_pooled + n] = fmaxf(m8[q] + bn, 0.f);
        }
    }
    __syncwarp();
}

// NW worker warps (gather / convert / epilogues) + producer + MMA issuer.  NW = 8: one CTA per SM; NW = 4: 192-thread
// CTAs, two per SM when a CTA needs at most half of the shared memory and 256 TMEM columns -- two tiles in flight per
// SM, each CTA's serial gather -> MMA -> epilogue -> MMA -> pool chain overlapping the other's.
template <int NPASS, int NW>
__global__ void __launch_bounds__((NW + 2) * 32, NW == 4 ? 2 : 1) sa_mlp2_kernel(Args a) {
    using C = Cfg<NPASS>;
    constexpr int IM = C::IMAGES;
    constexpr int NTHREADS = (NW + 2) * 32;
    constexpr int HALVES = NW / 4;          // worker warps per TMEM lane quarter: they split the 32-column groups
    constexpr int RSTEP = 2 * NW;           // rows covered by one gather pass (16 lanes per row)
    constexpr int RP = 128 / RSTEP;         // rows of a tile per worker thread
    const int NST = a.nst;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const Layout<NPASS> L(a.c1, a.c2, a.nst, a.slot_bytes, a.bias_smem);
    uint8_t *ring = base + L.ring();
    uint8_t *abuf = base + L.abuf();
    float *stage_all = reinterpret_cast<float *>(abuf);
    float *sb1 = reinterpret_cast<float *>(base + L.bias());
    float *sb2 = sb1 + 384;
    const float *b1p = a.bias_smem ? sb1 : a.b1, *b2p = a.bias_smem ? sb2 : a.b2;   // valid for n < c2 / n < c3
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(base + L.bars());
    unsigned long long *full = bars, *empty = bars + MAX_NST, *a_ready = bars + 2 * MAX_NST, *dbar = bars + 2 * MAX_NST + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * MAX_NST + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int k1 = atoms_of(a.c1), k2 = atoms_of(a.c2);
    const int bn1 = round16(a.c2), bn2 = round16(a.c3);
    const int CR = a.chunk_rows;
    const int nh1 = (bn1 + CR - 1) / CR, nh2 = (bn2 + CR - 1) / CR;
    const int nper = k1 * nh1 + k2 * nh2;   // weight chunks per tile
    const int my_tiles = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(a_ready, NW);
        mbar_init(&dbar[0], 1);
        mbar_init(&dbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (a.bias_smem) {
        for (int i = tid; i < 512; i += NTHREADS) {
            if (i < 384) sb1[i] = i < a.c2 ? __ldg(a.b1 + i) : 0.f;
            sb2[i] = i < a.c3 ? __ldg(a.b2 + i) : 0.f;
        }
    }
    __syncthreads();
    if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t abuf_a = smem_u32(abuf);

    if (warp == 0) {
        // ---------------- weight producer ----------------
        if (lane == 0) {
            const long long total = (long long)my_tiles * nper;
            for (long long Lc = 0; Lc < total; ++Lc) {
                const int s = (int)(Lc % NST);
                if (Lc >= NST) mbar_wait(&empty[s], (uint32_t)((Lc / NST) + 1) & 1);
                int i = (int)(Lc % nper);
                const uint8_t *wp;
                int bn, c, nh, kat;
                if (i < k1 * nh1) { wp = a.W1p; bn = bn1; kat = k1; c = i / nh1; nh = i % nh1; }
                else { i -= k1 * nh1; wp = a.W2p; bn = bn2; kat = k2; c = i / nh2; nh = i % nh2; }
                // gp_gemm_pack layout: n-tiles of 256 columns, per tile and k-chunk one image [bn_tile x 128 B] (hi, lo)
                const int ncol = nh * CR, jt = ncol >> 8, within = ncol & 255;
                const int bnj = min(256, bn - 256 * jt);
                const uint32_t rows = (uint32_t)min(CR, bnj - within);
                const size_t img = (size_t)bnj * 128;
                const uint8_t *tile = wp + (size_t)jt * kat * IM * (256 * 128);
                mbar_arrive_expect_tx(&full[s], IM * rows * 128);
                for (int w = 0; w < IM; ++w)
                    bulk_g2s(ring + ((size_t)s * IM + w) * a.slot_bytes, tile + ((size_t)c * IM + w) * img + (size_t)within * 128,
                             rows * 128, &full[s]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t a_hi = abuf_a, a_lo = abuf_a + (NPASS == 3 ? L.natoms * ATOM : 0);
            long long consumed = 0;
            uint32_t a_phase = 0;
            auto gemm = [&](int kat, int nh_cnt, int bn, uint32_t dcol) {
                for (int c = 0; c < kat; ++c)
                    for (int nh = 0; nh < nh_cnt; ++nh) {
                        const int s = (int)(consumed % NST);
                        mbar_wait(&full[s], (uint32_t)(consumed / NST) & 1);
                        tc_fence_after();
                        const uint32_t b_hi = smem_u32(ring + (size_t)s * IM * a.slot_bytes);
                        const uint32_t b_lo = b_hi + (NPASS == 3 ? a.slot_bytes : 0);
                        const int ncol = nh * CR;
                        const uint32_t idesc = make_idesc_bf16(128, (uint32_t)min(CR, min(256, bn - (ncol & ~255)) - (ncol & 255)));
                        const uint32_t d = tmem + dcol + ncol;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t ao = c * ATOM + kk * 32, bo = kk * 32;
                            umma_bf16(d, make_desc(a_hi + ao), make_desc(b_hi + bo), idesc, (c | kk) ? 1u : 0u);
                            if (NPASS == 3) {
                                umma_bf16(d, make_desc(a_lo + ao), make_desc(b_hi + bo), idesc, 1u);
                                umma_bf16(d, make_desc(a_hi + ao), make_desc(b_lo + bo), idesc, 1u);
                            }
                        }
                        umma_commit(&empty[s]);
                        ++consumed;
                    }
            };
            for (int t = 0; t < my_tiles; ++t) {
                mbar_wait(a_ready, a_phase); a_phase ^= 1;   // gathered rows are in the A buffer
                tc_fence_after();
                gemm(k1, nh1, bn1, 0);
                umma_commit(&dbar[0]);
                mbar_wait(a_ready, a_phase); a_phase ^= 1;   // hidden activations are in the A buffer
                tc_fence_after();
                gemm(k2, nh2, bn2, (uint32_t)a.d2col);
                umma_commit(&dbar[1]);
            }
        }
        __syncwarp();
    } else {
        // ---------------- gather / convert / epilogues ----------------
        const int w = tid - 64;                    // 0..32 NW - 1
        const int e = warp - 2;
        const int quarter = warp & 3;              // TMEM lane quarter of this warp
        const int half = e >> 2;                   // which 32-column groups (odd / even)
        const int row = 32 * quarter + lane;       // epilogue row
        const uint32_t lane_addr = tmem + ((uint32_t)(32 * quarter) << 16);
        const uint32_t A_hi = abuf_a, A_lo = abuf_a + (NPASS == 3 ? L.natoms * ATOM : 0);
        const int jv = w & 15;                     // float4 inside a 64-float chunk
        const int rsub = w >> 4;                   // row inside a pass
        const int ns = a.pool_ns;
        // Rows rsub + RSTEP p (p < RP) of a tile belong to this thread's gather.  Their ball-query indices for tile
        // t + 1 are requested while tile t is in its epilogues and only consumed at the next gather: the index load is
        // the head of the gather's latency chain.
        int gi[RP];
        auto request = [&](int t) {
            const long long row0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * BM;
#pragma unroll
            for (int p = 0; p < RP; ++p) {
                const long long gr = row0 + rsub + RSTEP * p;
                gi[p] = -1;
                if (t < my_tiles && gr < a.R) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(gi[p]) : "l"(a.gidx + gr));
            }
        };
        request(0);
        long long cy[6] = {0, 0, 0, 0, 0, 0};
        long long tt = clock64();
        auto lap = [&](int i) { const long long n = clock64(); cy[i] += n - tt; tt = n; };
        for (int t = 0; t < my_tiles; ++t) {
            const long long row0 = ((long long)blockIdx.x + (long long)t * gridDim.x) * BM;
            const uint32_t dph = (uint32_t)t & 1;
            // ---- gather: thread handles rows rsub + RSTEP p (p < RP), 16 bytes of every 64-float chunk, 8 rows at a time ----
            const int tile_batch = (int)(row0 / a.rows_per_batch);
#pragma unroll 1
            for (int pb = 0; pb < RP; pb += 8) {
                long long src[8];
                int qoff[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const int g32 = (int)(row0 + rsub + RSTEP * (pb + p));
                    const int batch = a.tile_in_batch ? tile_batch : g32 / a.rows_per_batch;
                    const int qrow = a.q_shift >= 0 ? g32 >> a.q_shift : g32 / a.q_ns;
                    const int gidx = gi[pb + p];
                    src[p] = gidx >= 0 ? ((long long)batch * a.n_src + gidx) * a.ldp : -1;
                    qoff[p] = gidx >= 0 ? qrow * a.ldq : 0;
                }
                for (int kc = 0; kc < k1; ++kc) {
                    const int k = kc * 64 + 4 * jv;
                    const bool k_ok = k < a.c1;
                    float4 pv[8], qv[8];
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const bool ok = k_ok && src[p] >= 0;
                        pv[p] = ok ? __ldg(reinterpret_cast<const float4 *>(a.P + src[p] + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        qv[p] = ok ? __ldg(reinterpret_cast<const float4 *>(a.Q + qoff[p] + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int p = 0; p < 8; ++p) {
                        const int rl = rsub + RSTEP * (pb + p);
                        const uint32_t off = kc * ATOM + rl * 128 + (((jv >> 1) ^ (rl & 7)) << 4) + ((jv & 1) << 3);
                        const float x0 = fmaxf(pv[p].x - qv[p].x, 0.f), x1 = fmaxf(pv[p].y - qv[p].y, 0.f);
                        const float x2 = fmaxf(pv[p].z - qv[p].z, 0.f), x3 = fmaxf(pv[p].w - qv[p].w, 0.f);
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(x0, x1), h1 = __floats2bfloat162_rn(x2, x3);
                        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(A_hi + off), "r"(*reinterpret_cast<const uint32_t *>(&h0)),
                                     "r"(*reinterpret_cast<const uint32_t *>(&h1)));
                        if (NPASS == 3) {
                            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
                            const __nv_bfloat162 l0 = __floats2bfloat162_rn(x0 - f0.x, x1 - f0.y);
                            const __nv_bfloat162 l1 = __floats2bfloat162_rn(x2 - f1.x, x3 - f1.y);
                            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(A_lo + off), "r"(*reinterpret_cast<const uint32_t *>(&l0)),
                                         "r"(*reinterpret_cast<const uint32_t *>(&l1)));
                        }
                    }
                }
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            lap(0);
            request(t + 1);
            lap(1);

            // ---- hidden layer: H = relu(D1 + b1) -> A buffer (zero beyond c2, up to whole k-atoms) ----
            mbar_wait(&dbar[0], dph);
            tc_fence_after();
            lap(2);
            for (int g = half; g < 2 * k2; g += HALVES) {
                uint32_t r[32];
                tmem_ld32(lane_addr + g * 32, r);
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    const int n0 = g * 32 + j8 * 8;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = n0 + j < a.c2 ? fmaxf(__uint_as_float(r[j8 * 8 + j]) + b1p[n0 + j], 0.f) : 0.f;
                    const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                    const __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                    uint4 pk;
                    pk.x = *reinterpret_cast<const uint32_t *>(&p0); pk.y = *reinterpret_cast<const uint32_t *>(&p1);
                    pk.z = *reinterpret_cast<const uint32_t *>(&p2); pk.w = *reinterpret_cast<const uint32_t *>(&p3);
                    const uint32_t off = (n0 >> 6) * ATOM + row * 128 + ((((n0 & 63) >> 3) ^ (row & 7)) << 4);
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
            if (lane == 0) mbar_arrive(a_ready);

            // ---- output layer + max-pool over the pool_ns rows of each group ----
            lap(3);
            mbar_wait(&dbar[1], dph);
            tc_fence_after();
            lap(4);
            const long long grow = row0 + row;
            const bool row_ok = grow < a.R;
            for (int g = half; g * 32 < a.c3; g += HALVES) {
                uint32_t r[32];
                tmem_ld32(lane_addr + a.d2col + g * 32, r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = row_ok ? __uint_as_float(r[j]) : -INFINITY;
                pool_block(v, stage_all + e * 1024, lane, row0 + 32 * quarter, a.R, ns, g * 32, a.c3, b2p, a.pooled, a.ld_pooled);
            }
            tc_fence_before();
            // the pooling stage lives in the A buffer: nobody gathers the next tile into it before all warps are done
            asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
            lap(5);
        }
#ifdef GP_SAF_PROBE
        if (blockIdx.x == 0 && tid == 64)
            printf("saf c=(%d,%d,%d) ns=%d tiles=%d per tile: gather %lld locate %lld wait_d0 %lld hidden %lld wait_d1 %lld pool %lld\n", a.c1, a.c2,
                   a.c3, a.pool_ns, my_tiles, cy[0] / my_tiles, cy[1] / my_tiles, cy[2] / my_tiles, cy[3] / my_tiles, cy[4] / my_tiles, cy[5] / my_tiles);
#endif
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)a.tmem_cols);
}

template <int NPASS>
static int launch(Args a, cudaStream_t st) {
    constexpr int IM = Cfg<NPASS>::IMAGES;
    const int bn1 = round16(a.c2), bn2 = round16(a.c3);
    // D2 reuses D1's TMEM columns (D1 is drained before the second GEMM starts)
    a.tmem_cols = (bn1 > 256 || bn2 > 256) ? 512 : 256;
    a.d2col = 0;
    a.tile_in_batch = a.rows_per_batch % BM == 0;
    a.q_shift = -1;
    for (int sh = 0; sh < 31; ++sh)
        if ((1 << sh) == a.q_ns) a.q_shift = sh;
    // One CTA per SM (its 10 warps are allocated as 12, two CTAs would leave 80 registers per thread).  The weight
    // ring wants >= 2 stages: 128-column chunks and biases in shared memory if that fits, else 64-column chunks,
    // else biases through L1.
    const size_t full_sm = 227 * 1024;
    int nst = 0;
    size_t fixed = 0, stage = 0;
    const int tries[3][2] = {{128, 1}, {64, 1}, {64, 0}};
    for (int t = 0; t < 3; ++t) {
        const int widest = bn1 > bn2 ? bn1 : bn2;
        a.chunk_rows = tries[t][0];
        a.bias_smem = tries[t][1];
        a.slot_bytes = (widest < a.chunk_rows ? widest : a.chunk_rows) * 128;
        fixed = Layout<NPASS>(a.c1, a.c2, 0, a.slot_bytes, a.bias_smem).total();
        stage = (size_t)IM * a.slot_bytes;
        nst = fixed + stage <= full_sm ? (int)((full_sm - fixed) / stage) : 0;
        if (nst >= 2) break;
    }
    GP_REQUIRE(nst >= 1, "gp_sa_mlp2_fused: layer widths need %zu bytes of shared memory", fixed + stage);
    a.nst = nst > MAX_NST ? MAX_NST : nst;
    // two 192-thread CTAs per SM when one of them (with at least one ring stage of 128-column chunks) fits in half
    // of the shared memory and 256 TMEM columns
    const size_t half_sm = 113 * 1024;
    if (a.tmem_cols == 256 && a.chunk_rows == 128 && fixed + stage <= half_sm) {
        const int nst2 = (int)((half_sm - fixed) / stage);
        a.nst = nst2 > MAX_NST ? MAX_NST : nst2;
        const Layout<NPASS> L2(a.c1, a.c2, a.nst, a.slot_bytes, a.bias_smem);
        auto kern = sa_mlp2_kernel<NPASS, 4>;
        GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        const int cap = 2 * num_sms();
        kern<<<a.ntiles < cap ? a.ntiles : cap, 6 * 32, L2.total(), st>>>(a);
        GP_CHECK_LAUNCH("gp_sa_mlp2_fused");
        return GP_OK;
    }
    const Layout<NPASS> L(a.c1, a.c2, a.nst, a.slot_bytes, a.bias_smem);
    auto kern = sa_mlp2_kernel<NPASS, 8>;
    const size_t smem = L.total();
    GP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    const int grid = a.ntiles < num_sms() ? a.ntiles : num_sms();
    kern<<<grid, 10 * 32, smem, st>>>(a);
    GP_CHECK_LAUNCH("gp_sa_mlp2_fused");
    return GP_OK;
}

}  // namespace saf
}  // namespace gp

using namespace gp;

extern "C" int gp_sa_mlp2_fused(const float *P, int n_src, int ldp, const int32_t *gidx, long long R, int rows_per_batch,
                                const float *Q, int ldq, int q_ns, const void *packed1, const float *bias1, int c1, int c2,
                                const void *packed2, const float *bias2, int c3, int npass, int pool_ns, float *pooled,
                                int ld_pooled, gp_stream_t s) {
    GP_REQUIRE(R >= 0 && (npass == 1 || npass == 3), "gp_sa_mlp2_fused: bad arguments");
    if (R == 0) return GP_OK;
    GP_REQUIRE(P && gidx && Q && packed1 && bias1 && packed2 && bias2 && pooled, "gp_sa_mlp2_fused: null pointer");
    GP_REQUIRE(c1 >= 4 && c1 <= 256 && (c1 & 3) == 0 && c2 >= 1 && c2 <= 384 && c3 >= 1 && c3 <= 512,
               "gp_sa_mlp2_fused: widths must satisfy c1 %% 4 == 0, c1 <= 256, c2 <= 384 and c3 <= 512 (got %d, %d, %d)", c1, c2, c3);
    GP_REQUIRE(ldp >= c1 && (ldp & 3) == 0 && ((uintptr_t)P & 15) == 0 && ldq >= c1 && (ldq & 3) == 0 && ((uintptr_t)Q & 15) == 0,
               "gp_sa_mlp2_fused: P / Q rows must be 16-byte aligned and cover c1");
    GP_REQUIRE(((uintptr_t)packed1 & 15) == 0 && ((uintptr_t)packed2 & 15) == 0, "gp_sa_mlp2_fused: packed weights must be 16-byte aligned");
    GP_REQUIRE(rows_per_batch >= 1 && n_src >= 1 && q_ns >= 1, "gp_sa_mlp2_fused: bad gather geometry");
    GP_REQUIRE((pool_ns == 8 || pool_ns == 16 || pool_ns == 32) && R % pool_ns == 0 && ld_pooled >= c3,
               "gp_sa_mlp2_fused: pool_ns must be 8, 16 or 32 and divide R");
    GP_REQUIRE(R < 2147483647LL && (R / q_ns + 1) * (long long)ldq < 2147483647LL, "gp_sa_mlp2_fused: too many rows");
    saf::Args a;
    a.P = P; a.n_src = n_src; a.ldp = ldp; a.gidx = gidx; a.R = R; a.rows_per_batch = rows_per_batch;
    a.Q = Q; a.ldq = ldq; a.q_ns = q_ns;
    a.W1p = (const uint8_t *)packed1; a.b1 = bias1; a.c1 = c1; a.c2 = c2;
    a.W2p = (const uint8_t *)packed2; a.b2 = bias2; a.c3 = c3;
    a.pool_ns = pool_ns; a.pooled = pooled; a.ld_pooled = ld_pooled;
    a.ntiles = (int)((R + saf::BM - 1) / saf::BM);
    return npass == 3 ? saf::launch<3>(a, as_stream(s)) : saf::launch<1>(a, as_stream(s));
}
