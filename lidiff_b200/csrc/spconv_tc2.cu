// K4 (variant B, persistent) — the output-stationary tcgen05 sparse convolution of spconv_tc.cu restructured as a
// persistent kernel: one CTA per SM walks the (pass, 128-row tile) work items round-robin and keeps every pipeline
// running ACROSS tiles, so the latency chain of a tile (neighbour-index fetch -> first gather -> MMA -> drain ->
// epilogue) overlaps with its neighbours instead of being paid ~20 times per SM and launch (measured: 22 us per tile
// on the 96-channel level-0 layers against a 5 us bandwidth floor).
//
//   warps 0-3  A producers: per tile fetch row ids (execution order) + the K neighbour rows of each, publish the
//              tile's non-empty-offset mask (meta_full), then stream the gathered rows of every (offset, chunk)
//              through the stage ring (cp.async from the fp16 split companions, or fp32 -> split in registers)
//   warp 4     MMA issuer (FP16x3, accumulator groups of <= STEP_BUDGET chained steps, ping-pong TMEM when it fits)
//   warp 5     weight loader (one cp.async.bulk pair per stage)
//   warps 6-9  drain: two-level accumulation (RN fp32 running total in TMEM) and the fused epilogue, coalesced
//              through a per-warp 32x32 shared-memory slab; outputs fp32 and/or fp16 split companions
//
// Same math, same parameters and the same results as k_spconv_tc (tests compare them).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace tc2 {
using namespace tc;

constexpr int THREADS = 320;
constexpr int MAX_KVOL = 27;
constexpr int MAX_STAGES = 4;
constexpr int SLAB_PITCH = 36;                        // floats per slab row (32 + 4: conflict-free 16-byte accesses)
constexpr int SLAB_BYTES = BM * SLAB_PITCH * 4;       // 4 warps x 32 rows
constexpr int META = 4;                               // ring of per-tile metadata (row ids, offset masks).  Must exceed the cp.async
                                                      // lookahead D <= MAX_STAGES-1: a tile's last full_a arrival is issued up to D
                                                      // stage-iterations (= up to D tiles) later, while re-using a slot waits for the
                                                      // tile META positions back to be completely drained.

struct Params {
    int c1, c2, cout, kvol;
    const unsigned char* wpacked;
    const float* scale;
    const float* shift;
    int relu;
    const int* nbr;
    long long nbr_stride;
    const int* d_mout;
    int mout_cap;
    const int* row_perm;
    const unsigned* row_mask;
    int stages, lag, nchunks, tmem_cols, tot_col, group, nbuf, acc_stride, npass;
    lb2_conv_io io[2];
};

__global__ void __launch_bounds__(THREADS, 1) k_spconv_tc_persist(const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const int M = p.d_mout ? min(*p.d_mout, p.mout_cap) : p.mout_cap;
    const int n_tiles = (M + BM - 1) / BM;
    const int total = n_tiles * p.npass;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctot = p.c1 + p.c2;

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t b_tile = (uint32_t)p.cout * 128u;
    const uint32_t stage_bytes = 2u * A_TILE + 2u * b_tile;
    unsigned char* tail = gen + (size_t)p.stages * stage_bytes;
    float* slab = reinterpret_cast<float*>(tail);                                   // [4 warps][32][SLAB_PITCH]
    int* idx_s = reinterpret_cast<int*>(tail + SLAB_BYTES);                          // [kvol][BM] (producer private)
    int* row_s = idx_s + MAX_KVOL * BM;                                              // [META][BM]
    uint32_t* wmask = reinterpret_cast<uint32_t*>(row_s + META * BM);                // [META][4] per-warp offset masks
    uint64_t* bars = reinterpret_cast<uint64_t*>(wmask + 4 * META);
    uint32_t* misc = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + 4 + 2 * META);
    const uint32_t bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto full_b = [&](int s) { return bar0 + 8u * (MAX_STAGES + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * MAX_STAGES + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (3 * MAX_STAGES + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (3 * MAX_STAGES + 2 + b); };
    auto meta_full = [&](int b) { return bar0 + 8u * (3 * MAX_STAGES + 4 + b); };
    auto meta_empty = [&](int b) { return bar0 + 8u * (3 * MAX_STAGES + 4 + META + b); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_a(s), 128); mbar_init(full_b(s), 1); mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 128); }
        for (int b = 0; b < META; ++b) { mbar_init(meta_full(b), 1); mbar_init(meta_empty(b), 130); }   // MMA + loader + 128 drain threads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc[0])), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = misc[0];
    const float out_scale = __ldg(reinterpret_cast<const float*>(p.wpacked) + 1);

    auto tile_kmask = [&](int b) { return wmask[b * 4] | wmask[b * 4 + 1] | wmask[b * 4 + 2] | wmask[b * 4 + 3]; };
    struct Ring {                                        // position in the stage ring without integer division
        int s; uint32_t par; int n;
        __device__ __forceinline__ void next() { if (++s == n) { s = 0; par ^= 1u; } }
    };

    if (warp < 4) {
        // =========================== A producers ===========================
        const int t = threadIdx.x;
        const int sub = t & 7, rbase = t >> 3;
        // cp.async lookahead.  A stage's full_a arrival is signalled D stage-iterations after its copies were issued; issuing
        // iteration it+D needs the slot of it+D-S free.  D = S-1 would couple "stage `it` may be consumed" to "the MMAs of stage
        // it-1 have retired" and leave the tensor pipe idle for the producers' reaction time once per stage; D = S-2 keeps one
        // full stage of slack.
        const int D = p.lag;
        int it = 0, arrived = 0, j = 0;
        Ring ri{0, 0u, p.stages}, ra{0, 0u, p.stages};                  // issue position / arrival position
        auto fetch_row = [&](int item) {                                // output row of this thread's slot in work item `item`
            if (item >= total) return -1;
            const int tile = (item >= n_tiles) ? item - n_tiles : item;
            const int slot = tile * BM + t;
            return (slot < M) ? (p.row_perm ? __ldg(p.row_perm + slot) : slot) : -1;
        };
        auto fetch_mask = [&](int row) -> uint32_t {                   // candidate offsets of a row: its neighbour bit mask if the caller has one
            if (row < 0) return 0u;
            return p.row_mask ? __ldg(p.row_mask + row) : ((p.kvol >= 32) ? 0xffffffffu : ((1u << p.kvol) - 1u));
        };
        int next_row = fetch_row(blockIdx.x);
        int next2_row = fetch_row(blockIdx.x + gridDim.x);
        uint32_t next_mask = fetch_mask(next_row);
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++j) {
            const int b = j % META;
            const int pass = (item >= n_tiles) ? 1 : 0;
            const lb2_conv_io io = p.io[pass];
            if (j >= META) mbar_wait(meta_empty(b), ((j / META) - 1) & 1);
            asm volatile("bar.sync 2, 128;" ::: "memory");              // everybody is done reading the previous tile's idx_s
            {
                const int row = next_row;
                const uint32_t have = next_mask;                        // offsets this row may have a neighbour at
                next_row = next2_row;
                next2_row = fetch_row(item + 2 * gridDim.x);            // prefetch two tiles ahead (row), one tile ahead (its mask):
                next_mask = fetch_mask(next_row);                       // their latency hides behind this tile's gathers
                row_s[b * BM + t] = row;
                uint32_t found = 0;
                for (int k0 = 0; k0 < p.kvol; k0 += 9) {                // up to 9 independent loads in flight, only for present offsets
                    int v[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int k = k0 + q;
                        v[q] = -1;
                        if (k < p.kvol && ((have >> k) & 1u)) v[q] = p.nbr ? __ldg(p.nbr + (long long)k * p.nbr_stride + row) : row;
                    }
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const int k = k0 + q;
                        if (k < p.kvol) {
                            idx_s[k * BM + t] = v[q];
                            if (v[q] >= 0) found |= 1u << k;
                        }
                    }
                }
                const uint32_t wm = __reduce_or_sync(0xffffffffu, p.row_mask ? have : found);
                if (lane == 0) wmask[b * 4 + warp] = wm;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (t == 0) mbar_arrive(meta_full(b));
            const uint32_t kmask = tile_kmask(b);
            const bool use_h = (io.in1_h != nullptr) && (p.c2 == 0 || io.in2_h != nullptr);
            for (uint32_t km = kmask; km; km &= km - 1) {
                const int k = __ffs(km) - 1;
                const int* idxk = idx_s + k * BM;
                int src[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) src[q] = idxk[rbase + 16 * q];
                for (int c = 0; c < p.nchunks; ++c, ++it, ri.next()) {
                    const int s = ri.s;
                    mbar_wait(empty(s), ri.par ^ 1u);
                    unsigned char* a_hi = gen + (size_t)s * stage_bytes;
                    const uint32_t a_hi_u = base + (uint32_t)s * stage_bytes;
                    const int ch = c * KC + sub * 8;
                    if (ch < ctot) {
                        const bool first = ch < p.c1;
                        const int cw = first ? p.c1 : p.c2;
                        const int co = first ? ch : ch - p.c1;
                        if (use_h) produce_a_split(reinterpret_cast<const __half*>(first ? io.in1_h : io.in2_h), cw, co, src, a_hi_u, a_hi_u + A_TILE, rbase, sub);
                        else produce_a_f32(first ? io.in1 : io.in2, cw, co, src, a_hi, a_hi + A_TILE, rbase, sub);
                    }
                    cp_async_commit();                                  // (empty group on the fp32 path)
                    if (it >= D) {
                        cp_async_wait_dyn(D);
                        fence_proxy_async();
                        mbar_arrive(full_a(ra.s));
                        ra.next();
                        ++arrived;
                    }
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (; arrived < it; ++arrived, ra.next()) mbar_arrive(full_a(ra.s));
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.cout);
            int gcount = 0, j = 0;
            Ring r{0, 0u, p.stages};
            const bool two = p.nbuf == 2;
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++j) {
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                const int n_off = __popc(kmask);
                int in_group = 0, off_idx = 0;
                for (uint32_t km = kmask; km; km &= km - 1, ++off_idx) {
                    const int buf = two ? (gcount & 1) : 0;
                    const int use = two ? (gcount >> 1) : gcount;         // how often this accumulator has been used before
                    const uint32_t tmem_acc = tmem_d + (uint32_t)(buf * p.acc_stride);
                    if (in_group == 0 && use >= 1) {
                        mbar_wait(acc_empty(buf), (use - 1) & 1);
                        tc_fence_after();
                    }
                    for (int c = 0; c < p.nchunks; ++c, r.next()) {
                        const int s = r.s;
                        const uint32_t par = r.par;
                        mbar_wait(full_b(s), par);
                        mbar_wait(full_a(s), par);
                        tc_fence_after();
                        const uint32_t a_hi = base + (uint32_t)s * stage_bytes, a_lo = a_hi + A_TILE;
                        const uint32_t b_hi = a_lo + A_TILE, b_lo = b_hi + b_tile;
                        const int ksteps = min(KC, ctot - c * KC) >> 4;
                        const uint64_t dah0 = make_desc(a_hi), dal0 = make_desc(a_lo), dbh0 = make_desc(b_hi), dbl0 = make_desc(b_lo);
                        for (int ks = 0; ks < ksteps; ++ks) {                 // +32 bytes per K step = +2 in the address field
                            const uint64_t dah = dah0 + 2u * ks, dal = dal0 + 2u * ks, dbh = dbh0 + 2u * ks, dbl = dbl0 + 2u * ks;
                            umma(tmem_acc, dah, dbh, idesc, (in_group | c | ks) ? 1u : 0u);
                            umma(tmem_acc, dal, dbh, idesc, 1);
                            umma(tmem_acc, dah, dbl, idesc, 1);
                        }
                        umma_commit(empty(s));
                    }
                    if (++in_group == p.group || off_idx == n_off - 1) {
                        umma_commit(acc_full(buf));
                        in_group = 0;
                        ++gcount;
                    }
                }
                mbar_arrive(meta_empty(b));
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // =========================== weight loader ===========================
        if (lane == 0) {
            int j = 0;
            Ring r{0, 0u, p.stages};
            for (int item = blockIdx.x; item < total; item += gridDim.x, ++j) {
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                for (uint32_t km = kmask; km; km &= km - 1) {
                    const int k = __ffs(km) - 1;
                    for (int c = 0; c < p.nchunks; ++c, r.next()) {
                        const int s = r.s;
                        mbar_wait(empty(s), r.par ^ 1u);
                        const uint32_t dst = base + (uint32_t)s * stage_bytes + 2u * A_TILE;
                        const unsigned char* src = p.wpacked + PACK_HEADER + ((size_t)k * p.nchunks + c) * (2u * b_tile);
                        mbar_expect_tx(full_b(s), 2u * b_tile);
                        bulk_g2s(dst, src, 2u * b_tile, full_b(s));
                    }
                }
                mbar_arrive(meta_empty(b));
            }
        }
        __syncwarp();
    } else {
        // =========================== drain + epilogue ===========================
        const int q4 = warp & 3;
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        float* myslab = slab + (size_t)q4 * 32 * SLAB_PITCH;
        int gcount = 0, j = 0;
        for (int item = blockIdx.x; item < total; item += gridDim.x, ++j) {
            const int b = j % META;
            const int pass = (item >= n_tiles) ? 1 : 0;
            const lb2_conv_io io = p.io[pass];
            mbar_wait(meta_full(b), (j / META) & 1);
            const uint32_t kmask = tile_kmask(b);
            const int n_off = __popc(kmask);
            const int n_groups = (n_off + p.group - 1) / p.group;
            const int* rows = row_s + b * BM + q4 * 32;                 // this warp's 32 output rows
            int orows[8], gidx[8];                                      // the 8 rows this lane serves in the epilogue + their gate rows
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                orows[i] = rows[(lane >> 3) + 4 * i];
                gidx[i] = (io.gate_table && io.gate_idx && orows[i] >= 0) ? __ldg(io.gate_idx + orows[i]) : 0;
            }
            for (int g = 0; g < max(n_groups, 1); ++g) {
                const bool last = g >= n_groups - 1;
                const int buf = (p.nbuf == 2) ? (gcount & 1) : 0;
                const uint32_t acc_col = (uint32_t)(buf * p.acc_stride);
                if (n_groups > 0) {
                    mbar_wait(acc_full(buf), ((p.nbuf == 2) ? (gcount >> 1) : gcount) & 1);
                    tc_fence_after();
                }
                for (int c0 = 0; c0 < p.cout; c0 += 32) {
                    float acc[32];
                    if (n_groups > 0) {
                        uint32_t r[32];
                        tmem_ld32(tmem_d + lane_base + acc_col + (uint32_t)c0, r);
                        if (g > 0) {
                            uint32_t tt[32];
                            tmem_ld32(tmem_d + lane_base + (uint32_t)(p.tot_col + c0), tt);
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 32; ++q) acc[q] = __fadd_rn(__uint_as_float(tt[q]), __uint_as_float(r[q]));
                        } else {
                            tmem_ld_wait();
#pragma unroll
                            for (int q = 0; q < 32; ++q) acc[q] = __uint_as_float(r[q]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) acc[q] = 0.f;
                    }
                    if (!last) {
                        uint32_t tt[32];
#pragma unroll
                        for (int q = 0; q < 32; ++q) tt[q] = __float_as_uint(acc[q]);
                        tmem_st32(tmem_d + lane_base + (uint32_t)(p.tot_col + c0), tt);
                    } else {
                        // ---- epilogue of this 32-column slab: transpose through the warp's slab, then lanes run along channels
                        __syncwarp();
                        float* srow = myslab + lane * SLAB_PITCH;
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(srow + q * 4) = make_float4(acc[q * 4] * out_scale, acc[q * 4 + 1] * out_scale,
                                                                                   acc[q * 4 + 2] * out_scale, acc[q * 4 + 3] * out_scale);
                        __syncwarp();
                        // all global loads of this slab first (8 rows per lane: memory-level parallelism), then the math + stores
                        const int lc4 = (lane & 7) * 4;
                        const int col = c0 + lc4;
                        float4 pre[8], res[8], gat[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            pre[i] = make_float4(0.f, 0.f, 0.f, 0.f); res[i] = pre[i]; gat[i] = make_float4(1.f, 1.f, 1.f, 1.f);
                            if (orows[i] >= 0) {
                                const long long ro = (long long)orows[i] * p.cout + col;
                                if (io.pre_add) pre[i] = __ldg(reinterpret_cast<const float4*>(io.pre_add + ro));
                                res[i] = load_residual4(io.residual, io.residual_h, orows[i], p.cout, col);
                                if (io.gate_table) gat[i] = __ldg(reinterpret_cast<const float4*>(io.gate_table + (long long)gidx[i] * p.cout + col));
                            }
                        }
                        float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), h4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (p.scale) { s4 = __ldg(reinterpret_cast<const float4*>(p.scale + col)); h4 = __ldg(reinterpret_cast<const float4*>(p.shift + col)); }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int orow = orows[i];
                            if (orow < 0) continue;
                            const int rr = (lane >> 3) + 4 * i;
                            const long long ro = (long long)orow * p.cout + col;
                            const float4 a4 = *reinterpret_cast<const float4*>(myslab + rr * SLAB_PITCH + lc4);
                            float y[4] = {a4.x + pre[i].x, a4.y + pre[i].y, a4.z + pre[i].z, a4.w + pre[i].w};
                            y[0] = fmaf(y[0], s4.x, h4.x) + res[i].x; y[1] = fmaf(y[1], s4.y, h4.y) + res[i].y;
                            y[2] = fmaf(y[2], s4.z, h4.z) + res[i].z; y[3] = fmaf(y[3], s4.w, h4.w) + res[i].w;
                            if (p.relu) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) y[q] = fmaxf(y[q], 0.f);
                            }
                            if (io.out) *reinterpret_cast<float4*>(io.out + ro) = make_float4(y[0], y[1], y[2], y[3]);
                            if (io.out_h) store_split4(io.out_h, orow, p.cout, col, y);
                            if (io.out_gated || io.out_gated_h) {
                                y[0] *= gat[i].x; y[1] *= gat[i].y; y[2] *= gat[i].z; y[3] *= gat[i].w;
                                if (io.out_gated) *reinterpret_cast<float4*>(io.out_gated + ro) = make_float4(y[0], y[1], y[2], y[3]);
                                if (io.out_gated_h) store_split4(io.out_gated_h, orow, p.cout, col, y);
                            }
                        }
                    }
                }
                if (n_groups > 0) {
                    if (!last) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    mbar_arrive(acc_empty(buf));            // persistent: always release (the next tile reuses it)
                    ++gcount;
                }
            }
            mbar_arrive(meta_empty(b));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
}

static size_t smem_bytes(int cout, int stages) {
    return 1024 + (size_t)stages * (2 * A_TILE + 2 * (size_t)cout * 128) + SLAB_BYTES + MAX_KVOL * BM * sizeof(int) + META * BM * sizeof(int) +
           4 * META * sizeof(uint32_t) + (3 * MAX_STAGES + 4 + 2 * META) * 8 + 64;
}

}  // namespace tc2


bool lb2_spconv_tc3_supported(const lb2_conv_desc* d);
int lb2_spconv_tc3_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget);
bool lb2_spconv_tc4_supported(const lb2_conv_desc* d);
int lb2_spconv_tc4_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget);
bool lb2_spconv_tc5_supported(const lb2_conv_desc* d);
int lb2_spconv_tc5_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget);

int lb2_spconv_tc2_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget) {
    const int use_pair = h->opt[LB2_OPT_TC_PAIR];           // CTA-pair kernel (cta_group::2): 0 off, 1 = Cout 256 only, 2 = Cout 256 and 128
    if (use_pair && (d->cout == 256 || use_pair >= 2) && lb2_spconv_tc5_supported(d)) return lb2_spconv_tc5_launch(h, s, d, step_budget);
    if (h->opt[LB2_OPT_TC_N256] && lb2_spconv_tc3_supported(d)) return lb2_spconv_tc3_launch(h, s, d, step_budget);
    if (h->opt[LB2_OPT_TC_SMALL] && lb2_spconv_tc4_supported(d)) return lb2_spconv_tc4_launch(h, s, d, step_budget);
    tc2::Params p;
    p.c1 = d->c1; p.c2 = d->c2; p.cout = d->cout; p.kvol = d->kvol; p.npass = d->npass;
    p.wpacked = (const unsigned char*)d->weight_packed;
    p.scale = d->scale; p.shift = d->shift; p.relu = d->relu;
    p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.d_mout = d->d_mout; p.mout_cap = d->mout_cap; p.row_perm = d->row_perm; p.row_mask = d->row_mask;
    p.nchunks = (d->c1 + d->c2 + tc::KC - 1) / tc::KC;
    int stages = tc2::MAX_STAGES;
    while (stages > 1 && tc2::smem_bytes(d->cout, stages) > 227 * 1024) --stages;
    p.stages = stages;
    {   // LB2_TC_LAG=full restores the S-1 lookahead (development A/B knob)
        p.lag = (stages >= 3 && !h->opt[LB2_OPT_TC_FULL_LAG]) ? stages - 2 : stages - 1;
    }
    const int half = d->cout <= 32 ? 32 : d->cout <= 64 ? 64 : d->cout <= 128 ? 128 : 256;
    p.nbuf = half <= 128 ? 2 : 1;
    p.acc_stride = half;
    p.tot_col = p.nbuf * half;
    { int need = (p.nbuf + 1) * half; p.tmem_cols = 32; while (p.tmem_cols < need) p.tmem_cols <<= 1; }
    const int steps_per_offset = 3 * ((d->c1 + d->c2 + 15) / 16);
    p.group = std::max(1, step_budget / steps_per_offset);
    p.io[0] = d->io[0]; p.io[1] = d->io[d->npass > 1 ? 1 : 0];
    const size_t smem = tc2::smem_bytes(d->cout, stages);
    {
        cudaError_t e = lb2_configure_smem(h, LB2_K_TC2, tc2::k_spconv_tc_persist, (int)(227 * 1024));
        if (e != cudaSuccess) return lb2_fail(h, LB2_ERR_CUDA, "k_spconv_tc_persist smem attribute: %s", cudaGetErrorString(e));
    }
    const long long tiles_cap = (long long)cdiv(d->mout_cap, tc::BM) * d->npass;
    const unsigned grid = (unsigned)std::min<long long>(h->num_sms, tiles_cap);
    tc2::k_spconv_tc_persist<<<grid, tc2::THREADS, smem, s>>>(p);
    LB2_POST_LAUNCH(h, "k_spconv_tc_persist");
    return LB2_OK;
}
