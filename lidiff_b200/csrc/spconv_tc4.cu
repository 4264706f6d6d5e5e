// K4 (variant B, persistent, Cout <= 128) — the persistent tcgen05 sparse convolution of spconv_tc2.cu re-balanced for the
// layers that are bound by memory latency and by the epilogue rather than by the tensor pipe (levels 0-2, level-3 encoder).
// Profiling the generic kernel on those layers (profiles/r01_ncu_stall_attribution.txt) showed its four drain warps busy
// >95 % of the time (one warp per scheduler, a long dependent instruction stream per tile) while the MMA warp waited for
// a free accumulator and the producers for a free stage.  This kernel therefore
//   * runs TWO drain warpgroups (alternate tiles for Cout <= 64, half the channels each for Cout 96 / 128), keeps the fp32 running total of the two-level accumulation in
//     their registers (setmaxnreg: 96 / 48 / 184 / 184 registers for producer / MMA+loader / drain / drain warpgroups) and
//     releases an accumulator as soon as it has been read, so TMEM holds four accumulators the MMA warp can run ahead into;
//   * separates the gathered-A ring (4 slots of 32 K-columns: the two halves of a 128-byte swizzled row image are disjoint
//     sets of 16-byte chunks) from the weight ring (3 slots of 64 K-columns), with a gather lookahead of two slots.
//   WG0 warps 0-3   A producers (cp.async from the fp16 split companions, or fp32 -> split in registers)
//   WG1 warp 4 MMA issuer, warp 5 weight loader (warps 6,7 idle)
//   WG2 warps 8-11  drain + epilogue of even tiles / the low channels        WG3 warps 12-15: odd tiles / the high channels
// Same math and results as k_spconv_tc / k_spconv_tc_persist (tests compare them).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace tc4 {
using namespace tc;

constexpr int THREADS = 512;
constexpr int MAX_KVOL = 27;
constexpr int NA_MAX = 6;
__host__ __device__ constexpr int na_of(int ncc) { return ncc <= 3 ? 6 : 4; }   // A slots (32 K-columns each; two per 128-byte row image):
                                                      // the sparse levels' layers (Cout <= 96) are bound by gather latency, their
                                                      // smaller weight slots leave room for a deeper gather ring
constexpr int NB = 3;                                 // weight slots (64 K-columns each)
constexpr int NACC = 4;                               // TMEM accumulators
constexpr int SLAB_PITCH = tc::EPI_PITCH;             // floats per slab row
constexpr int SLAB_BYTES = 8 * 32 * SLAB_PITCH * 4;   // 8 drain warps x 32 rows
constexpr int META = 4;                               // ring of per-tile metadata; must exceed the gather lookahead in tiles (<= 1)
constexpr int NBAR = 2 * NA_MAX + 2 * NB + 2 * NACC + 2 * META;

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
    int nchunks, nhalf, tmem_cols, group, acc_stride, npass;
    const int* tile_order;      // tiles by descending cost (lb2_tile_order) or NULL
    lb2_conv_io io[2];
};

struct Ring {                                             // position in a ring without integer division
    int s; uint32_t par; int n;
    __device__ __forceinline__ void next() { if (++s == n) { s = 0; par ^= 1u; } }
};

template <int NCC>                                        // NCC = Cout / 32: sizes the register-resident running total
__global__ void __launch_bounds__(THREADS, 1) k_spconv_tc_small(const Params p) {
    extern __shared__ unsigned char smem_raw[];
    constexpr int NA = na_of(NCC);
    const int M = p.d_mout ? min(*p.d_mout, p.mout_cap) : p.mout_cap;
    const int n_tiles = (M + BM - 1) / BM;
    const int total = n_tiles * p.npass;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item -> (tile, pass).  With a cost order (lb2_tile_order): item i = pass (i & 1) of the (i >> 1)-th most expensive tile, so the
    // item sequence is sorted by cost and the snake deal below is balanced to 1-3 %.  Without: pass-major, inside a pass the last
    // tiles of the mask-sorted row order (roughly the heaviest) first.
    const int pshift = (p.tile_order && p.npass == 2) ? 1 : 0;
    auto item_pass = [&](int item) { return p.tile_order ? (item & pshift) : (item >= n_tiles ? 1 : 0); };
    auto item_tile = [&](int item) {
        if (p.tile_order) return __ldg(p.tile_order + (item >> pshift));
        return n_tiles - 1 - (item >= n_tiles ? item - n_tiles : item);
    };
    // round j of the persistent loop in snake order (even rounds left to right, odd rounds right to left over the CTAs): with the
    // items sorted by cost every CTA alternates between a dearer and a cheaper item, so the per-CTA sums stay balanced (static LPT);
    // an item index >= total (last, partial round) is an empty tile for every role
    auto slot_item = [&](int jj) { return jj * (int)gridDim.x + ((jj & 1) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x); };

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t b_tile = (uint32_t)p.cout * 128u;
    const uint32_t a_stage = 2u * A_TILE;                       // hi + lo image of 128 rows x 64 K-columns (= 2 A slots)
    const uint32_t b_base = base + (NA / 2) * a_stage;
    unsigned char* tail = gen + (size_t)(NA / 2) * a_stage + (size_t)NB * 2u * b_tile;
    float* slab = reinterpret_cast<float*>(tail);                                   // [8 warps][32][SLAB_PITCH]
    int* idx_s = reinterpret_cast<int*>(tail + SLAB_BYTES);                          // [kvol][BM] (producer private)
    int* row_s = idx_s + MAX_KVOL * BM;                                              // [META][BM]
    int* gate_s = row_s + META * BM;                                                 // [8 drain warps][32] gate-table rows of the tile being drained
    uint32_t* wmask = reinterpret_cast<uint32_t*>(gate_s + 8 * 32);                  // [META][4] per-warp offset masks
    uint64_t* bars = reinterpret_cast<uint64_t*>(wmask + 4 * META);
    uint32_t* misc = reinterpret_cast<uint32_t*>(bars + NBAR);
    float* aff_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(misc + 16) + 15) & ~uintptr_t(15));                              // [2][cout] BN scale, shift
    const uint32_t bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto empty_a = [&](int s) { return bar0 + 8u * (NA + s); };
    auto full_b = [&](int s) { return bar0 + 8u * (2 * NA + s); };
    auto empty_b = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + NACC + b); };
    auto meta_full = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + 2 * NACC + b); };
    auto meta_empty = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + 2 * NACC + META + b); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(full_a(s), 128); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int b = 0; b < NACC; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), NCC >= 3 ? 256 : 128); }    // the drain warpgroup(s) reading it
        for (int b = 0; b < META; ++b) { mbar_init(meta_full(b), 1); mbar_init(meta_empty(b), 258); }   // MMA + loader + both drain warpgroups
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_affine(aff_s, p.scale, p.shift, p.cout);
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

    if (warp < 4) {
        // =========================== WG0: A producers ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
        const int t = threadIdx.x;
        const int sub = t & 3, rbase = t >> 2;                          // 16-byte chunk inside the half row / first of this thread's 4 rows
        int j = 0;
        Ring ri{0, 0u, NA};                                             // issue position
        auto fetch_row = [&](int item) {                                // output row of this thread's slot in work item `item`
            if (item >= total) return -1;
            const int slot = item_tile(item) * BM + t;
            return (slot < M) ? (p.row_perm ? __ldg(p.row_perm + slot) : slot) : -1;
        };
        auto fetch_mask = [&](int row) -> uint32_t {                   // candidate offsets of a row: its neighbour bit mask if the caller has one
            if (row < 0) return 0u;
            return p.row_mask ? __ldg(p.row_mask + row) : ((p.kvol >= 32) ? 0xffffffffu : ((1u << p.kvol) - 1u));
        };
        int next_row = fetch_row(slot_item(0));
        int next2_row = fetch_row(slot_item(1));
        uint32_t next_mask = fetch_mask(next_row);
        for (; j * (int)gridDim.x < total; ++j) {
            const int item = slot_item(j);
            const int b = j % META;
            const lb2_conv_io io = p.io[item_pass(item)];
            if (j >= META) mbar_wait(meta_empty(b), ((j / META) - 1) & 1);
            asm volatile("bar.sync 2, 128;" ::: "memory");              // everybody is done reading the previous tile's idx_s
            {
                const int row = next_row;
                const uint32_t have = next_mask;                        // offsets this row may have a neighbour at
                next_row = next2_row;
                next2_row = fetch_row(slot_item(j + 2));            // prefetch two tiles ahead (row), one tile ahead (its mask):
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
                int src[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) src[q] = idx_s[k * BM + rbase + 32 * q];
                for (int c2 = 0; c2 < p.nhalf; ++c2, ri.next()) {
                    const int s = ri.s;
                    mbar_wait(empty_a(s), ri.par ^ 1u);
                    const uint32_t img = (uint32_t)(s >> 1) * a_stage;
                    const int ch = c2 * 32 + sub * 8;                   // first of this thread's 8 input channels
                    const bool first = ch < p.c1;
                    const int cw = first ? p.c1 : p.c2;
                    const int co = first ? ch : ch - p.c1;
                    if (use_h) {
                        const __half* src_h = reinterpret_cast<const __half*>(first ? io.in1_h : io.in2_h);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t off = img + sw128(rbase + 32 * q, (s & 1) * 4 + sub);
                            const bool ok = src[q] >= 0;
                            const __half* rp = src_h + (ok ? ((long long)src[q] * 2 * cw + co) : 0);
                            cp_async16(base + off, rp, ok ? 16u : 0u);
                            cp_async16(base + A_TILE + off, rp + (ok ? cw : 0), ok ? 16u : 0u);
                        }
                        cp_async_arrive_on(full_a(s));                  // published by the hardware when this thread's copies have landed
                    } else {
                        const float* srcp = first ? io.in1 : io.in2;
                        float4 va[4], vb[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            va[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                            vb[q] = va[q];
                            if (src[q] >= 0) {
                                const float4* rp = reinterpret_cast<const float4*>(srcp + (long long)src[q] * cw + co);
                                va[q] = __ldg(rp);
                                vb[q] = __ldg(rp + 1);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint4 hi, lo;
                            split8(va[q], vb[q], hi, lo);
                            const uint32_t off = img + sw128(rbase + 32 * q, (s & 1) * 4 + sub);
                            *reinterpret_cast<uint4*>(gen + off) = hi;
                            *reinterpret_cast<uint4*>(gen + A_TILE + off) = lo;
                        }
                        fence_proxy_async();                            // generic-proxy stores -> visible to the tensor core's reads
                        mbar_arrive(full_a(s));
                    }
                }
            }
        }
    } else if (warp < 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
        if (warp == 4) {
            // =========================== MMA issuer (whole warp converged; one elected lane issues) ===========================
            const uint32_t idesc = make_idesc(p.cout);
            int gcount = 0, j = 0;
            Ring rq{0, 0u, NA}, rb{0, 0u, NB};
            for (; j * (int)gridDim.x < total; ++j) {
            const int item = slot_item(j);
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                const int n_off = __popc(kmask);
                int in_group = 0, off_idx = 0;
                for (uint32_t km = kmask; km; km &= km - 1, ++off_idx) {
                    const int buf = gcount & (NACC - 1);
                    const uint32_t tmem_acc = tmem_d + (uint32_t)(buf * p.acc_stride);
                    if (in_group == 0 && gcount >= NACC) {
                        mbar_wait(acc_empty(buf), ((gcount / NACC) - 1) & 1);
                        tc_fence_after();
                    }
                    for (int c = 0; c < p.nchunks; ++c, rb.next()) {
                        const int sb = rb.s;
                        mbar_wait(full_b(sb), rb.par);
                        const uint32_t b_hi = b_base + (uint32_t)sb * 2u * b_tile, b_lo = b_hi + b_tile;
                        const uint64_t dbh0 = make_desc(b_hi), dbl0 = make_desc(b_lo);
                        const int halves = min(2, p.nhalf - 2 * c);
                        for (int half = 0; half < halves; ++half, rq.next()) {
                            const int sa = rq.s;
                            mbar_wait(full_a(sa), rq.par);
                            tc_fence_after();
                            const uint32_t a_hi = base + (uint32_t)(sa >> 1) * a_stage, a_lo = a_hi + A_TILE;
                            const uint64_t dah0 = make_desc(a_hi) + 4u * (uint32_t)(sa & 1), dal0 = make_desc(a_lo) + 4u * (uint32_t)(sa & 1);
                            if (elect_one()) {
#pragma unroll
                                for (int k2 = 0; k2 < 2; ++k2) {          // +32 bytes per K step = +2 in the descriptor's address field
                                    const uint32_t kb = 2u * (uint32_t)(half * 2 + k2);
                                    const uint64_t dah = dah0 + 2u * k2, dal = dal0 + 2u * k2, dbh = dbh0 + kb, dbl = dbl0 + kb;
                                    umma(tmem_acc, dah, dbh, idesc, (in_group | c | half | k2) ? 1u : 0u);
                                    umma(tmem_acc, dal, dbh, idesc, 1);
                                    umma(tmem_acc, dah, dbl, idesc, 1);
                                }
                                umma_commit(empty_a(sa));
                            }
                            __syncwarp();
                        }
                        if (elect_one()) umma_commit(empty_b(sb));
                        __syncwarp();
                    }
                    if (++in_group == p.group || off_idx == n_off - 1) {
                        if (elect_one()) umma_commit(acc_full(buf));
                        __syncwarp();
                        in_group = 0;
                        ++gcount;
                    }
                }
                if (lane == 0) mbar_arrive(meta_empty(b));
                __syncwarp();
            }
        } else if (warp == 5 && lane == 0) {
            // =========================== weight loader ===========================
            int j = 0;
            Ring r{0, 0u, NB};
            for (; j * (int)gridDim.x < total; ++j) {
            const int item = slot_item(j);
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                for (uint32_t km = kmask; km; km &= km - 1) {
                    const int k = __ffs(km) - 1;
                    for (int c = 0; c < p.nchunks; ++c, r.next()) {
                        const int s = r.s;
                        mbar_wait(empty_b(s), r.par ^ 1u);
                        const uint32_t dst = b_base + (uint32_t)s * 2u * b_tile;
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
        // =========================== WG2 / WG3: drain (register-resident fp32 total) + epilogue ===========================
        // Cout <= 64: the warpgroups take alternate tiles (all channels); Cout 96 / 128: both take every tile, half the channels each
        // (48 / 64 running totals per thread either way leave the registers for a 4-row epilogue batch).
        asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
        constexpr bool COLSPLIT = NCC >= 3;
        constexpr int TOT = COLSPLIT ? 16 * NCC : 32 * NCC;       // channels this warp drains per tile
        const int q4 = warp & 3;                                   // TMEM lane quarter
        const int wg = (warp >= 12) ? 1 : 0;
        const int cb = COLSPLIT ? wg * TOT : 0;                    // first of them
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        float* myslab = slab + (size_t)(warp - 8) * 32 * SLAB_PITCH;
        float tot[TOT];
        int gcount = 0, j = 0;
        for (; j * (int)gridDim.x < total; ++j) {
            const int item = slot_item(j);
            const int b = j % META;
            const lb2_conv_io& io = p.io[item_pass(item)];                // fields are read from the parameter bank when used
            mbar_wait(meta_full(b), (j / META) & 1);
            const uint32_t kmask = tile_kmask(b);
            const int n_off = __popc(kmask);
            int n_groups = 0;
            for (int o = 0; o < n_off; o += p.group) ++n_groups;
            if (!COLSPLIT && (j & 1) != wg) {                       // the other warpgroup's tile: only keep the accumulator count in step
                gcount += n_groups;
                mbar_arrive(meta_empty(b));
                continue;
            }
            const int* rows = row_s + b * BM + q4 * 32;             // this warp's 32 output rows
            int orows[4], gidx[4];                                  // the 4 rows this lane serves in the epilogue, their gate-table rows
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                orows[i] = rows[(lane >> 2) + 8 * i];
                gidx[i] = (io.gate_table && io.gate_idx && orows[i] >= 0) ? __ldg(io.gate_idx + orows[i]) : 0;
            }
            const unsigned fl = epi_flags(io, p.relu);
            if (fl & EP_OPERANDS) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {                       // L2 prefetch of the epilogue operands of this lane's 4 rows
                    if (orows[i] < 0) continue;
                    prefetch_row_f32(io.residual, orows[i], p.cout, cb, TOT, lane & 3);
                    if (!io.residual) prefetch_row_split(io.residual_h, orows[i], p.cout, cb, TOT, lane & 3);
                    prefetch_row_f32(io.pre_add, orows[i], p.cout, cb, TOT, lane & 3);
                    if (io.gate_table && io.gate_idx) prefetch_row_f32(io.gate_table, gidx[i], p.cout, cb, TOT, lane & 3);
                }
            }
            if (n_groups == 0) {
#pragma unroll
                for (int q = 0; q < TOT; ++q) tot[q] = 0.f;
            }
            for (int g = 0; g < n_groups; ++g) {
                const int buf = gcount & (NACC - 1);
                mbar_wait(acc_full(buf), (gcount / NACC) & 1);
                tc_fence_after();
                drain_acc<TOT>(tmem_d + lane_base + (uint32_t)(buf * p.acc_stride + cb), tot, g == 0);
                tc_fence_before();
                mbar_arrive(acc_empty(buf));                       // accumulator free again: the MMA warp runs on while we finish
                ++gcount;
            }
            // ---- epilogue from registers, 16 channels at a time through the warp's slab (tc_common.cuh: epilogue_slabs) ----
            constexpr int RB = (TOT >= 64) ? 2 : 4;                 // rows per load batch; 2 where the totals fill the registers
            epilogue_slabs<TOT, RB>(tot, myslab, lane, orows, gidx, cb, p.cout, out_scale, fl, io, aff_s);
            mbar_arrive(meta_empty(b));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)p.tmem_cols) : "memory");
}

static size_t smem_bytes(int cout) {
    return 1024 + (size_t)(na_of(cout / 32) / 2) * 2 * A_TILE + (size_t)NB * 2 * cout * 128 + SLAB_BYTES + MAX_KVOL * BM * sizeof(int) + META * BM * sizeof(int) +
           8 * 32 * sizeof(int) + 4 * META * sizeof(uint32_t) + NBAR * 8 + 64 + 16 + 2 * cout * sizeof(float);
}

}  // namespace tc4

bool lb2_spconv_tc4_supported(const lb2_conv_desc* d) {
    if (d->cout > 128 || d->cout % 32 != 0) return false;
    if ((long long)d->mout_cap * 2 * d->cout >= (1LL << 32)) return false;   // the epilogue indexes rows with 32-bit element offsets
    if ((d->c1 + d->c2) % 32 != 0 || d->c1 % 32 != 0) return false;          // whole 32-column A slots; a slot never straddles in1/in2
    return tc4::smem_bytes(d->cout) <= 227 * 1024;
}

int lb2_spconv_tc4_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget) {
    tc4::Params p;
    p.c1 = d->c1; p.c2 = d->c2; p.cout = d->cout; p.kvol = d->kvol; p.npass = d->npass;
    p.wpacked = (const unsigned char*)d->weight_packed;
    p.scale = d->scale; p.shift = d->shift; p.relu = d->relu;
    p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.d_mout = d->d_mout; p.mout_cap = d->mout_cap; p.row_perm = d->row_perm; p.row_mask = d->row_mask;
    p.nchunks = (d->c1 + d->c2 + tc::KC - 1) / tc::KC;
    p.nhalf = (d->c1 + d->c2) / 32;
    p.acc_stride = d->cout <= 32 ? 32 : d->cout <= 64 ? 64 : 128;
    p.tmem_cols = tc4::NACC * p.acc_stride;                       // 128 / 256 / 512: powers of two >= 32
    const int steps_per_offset = 3 * ((d->c1 + d->c2 + 15) / 16);
    p.group = std::max(1, step_budget / steps_per_offset);
    p.io[0] = d->io[0]; p.io[1] = d->io[d->npass > 1 ? 1 : 0];
    p.tile_order = d->tile_order128;
    const size_t smem = tc4::smem_bytes(d->cout);
    if (!(h->configured & (1u << LB2_K_TC4))) {
        cudaError_t e = cudaFuncSetAttribute(tc4::k_spconv_tc_small<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc4::k_spconv_tc_small<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc4::k_spconv_tc_small<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tc4::k_spconv_tc_small<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e != cudaSuccess) return lb2_fail(h, LB2_ERR_CUDA, "k_spconv_tc_small smem attribute: %s", cudaGetErrorString(e));
        h->configured |= 1u << LB2_K_TC4;
    }
    const long long tiles_cap = (long long)cdiv(d->mout_cap, tc::BM) * d->npass;
    const unsigned grid = (unsigned)std::min<long long>(h->num_sms, tiles_cap);
    switch (d->cout / 32) {
        case 1: tc4::k_spconv_tc_small<1><<<grid, tc4::THREADS, smem, s>>>(p); break;
        case 2: tc4::k_spconv_tc_small<2><<<grid, tc4::THREADS, smem, s>>>(p); break;
        case 3: tc4::k_spconv_tc_small<3><<<grid, tc4::THREADS, smem, s>>>(p); break;
        default: tc4::k_spconv_tc_small<4><<<grid, tc4::THREADS, smem, s>>>(p); break;
    }
    LB2_POST_LAUNCH(h, "k_spconv_tc_small");
    return LB2_OK;
}
