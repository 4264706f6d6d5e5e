// K4 (variant C, persistent CTA pairs, Cout = 128 / 256) — the sparse convolution of spconv_tc3.cu / spconv_tc4.cu on a cluster
// of two CTAs driving ONE tcgen05.mma.cta_group::2 stream (UMMA M = 256).
//
// Why: the single-CTA kernels are bound by L2 -> SM traffic, not by the tensor pipe (profiles/r01_ncu_spconv_n256_full_summary.txt:
// tensor pipe 49-66 % active at 9.2 TB/s of L2 -> SM reads, two thirds of it weight tiles that every CTA re-streams for every
// (tile, offset)).  With cta_group::2 the pair works on a super-tile of 256 output rows; each CTA gathers the A rows of its own
// 128 output rows as before but stages only HALF of every weight tile (its 128 / 64 of the Cout columns): the tensor cores of
// both SMs read both halves through the pair's shared-memory window.  Weight bytes per SM halve (Cout 256: 384 -> 256 KB per
// (tile, offset) including the gathered rows), and the freed shared memory deepens the gather ring (6 slots).
// The pair executes the UNION of its two tiles' neighbour offsets; in the mask-sorted row order adjacent tiles have nearly the
// same offset set, so the issued MMA work grows by 3-4 % only (measured on the bench geometry).
//
//   per CTA (512 threads, setmaxnreg 56 / 56 / 200 / 200):
//   WG0 warps 0-3   A producers: cp.async 16 B from the fp16 split companions into the SWIZZLE_128B image; a slot is published by
//                   cp.async.mbarrier.arrive.noinc the moment its copies land (no wait_group in the issue loop: all 6 slots are lookahead)
//   WG1 warp 4      rank 0: MMA issuer (elected lane, cta_group::2, commits multicast to both CTAs' barriers)
//                   rank 1: relay — forwards "my weight half / my gathered slot has landed" to the leader's barriers
//       warp 5      weight loader: two bulk copies (hi / lo half tile) per slot into its own shared memory
//   WG2/WG3 warps 8-15: drain (two-level accumulation: fp32 running total in registers) + fused epilogue, each CTA for its own
//                   128 rows / TMEM lanes; accumulator release is signalled to the leader (one arrive per warp)
// Math and results: FP16x3 operand split, fp32 accumulation in chains of <= STEP_BUDGET MMA steps, round-to-nearest totals —
// identical to k_spconv_tc_n256 / k_spconv_tc_small (rows that lack an offset of the union contribute exact zeros).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace tc5 {
using namespace tc;

constexpr int THREADS = 512;
constexpr int NA = 6;                                 // A slots (128 rows x 32 K-columns, hi + lo image: 16 KB each)
constexpr int SLAB_PITCH = tc::EPI_PITCH;             // floats per slab row
constexpr int SLAB_BYTES = 8 * 32 * SLAB_PITCH * 4;   // 8 drain warps x 32 rows
constexpr int META = 4;                               // ring of per-tile metadata (row ids, masks); exceeds the gather lookahead in tiles

template <int NCOLS> struct Cfg {
    static constexpr int NB = (NCOLS == 256) ? 3 : 4;              // weight slots (64 K-columns, this CTA's half of the columns, hi + lo)
    static constexpr int NACC = 512 / NCOLS;                       // TMEM accumulators (ping-pong / 4-deep)
    static constexpr uint32_t B_HALF = (uint32_t)(NCOLS / 2) * 128u;   // bytes of this CTA's half of one hi (or lo) weight tile
    static constexpr uint32_t B_SLOT = 2u * B_HALF;
    static constexpr int NBAR = 3 * NA + 3 * NB + 2 * NACC + 2 * META;
    static constexpr size_t SMEM = 1024 + (size_t)(NA / 2) * 2 * A_TILE + (size_t)NB * B_SLOT + SLAB_BYTES + META * BM * sizeof(int) +
                                   META * BM * sizeof(uint32_t) + 4 * META * sizeof(uint32_t) + NBAR * 8 + 64 + 16 + 2 * NCOLS * sizeof(float);
};

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
    int nchunks, group, npass;
    const int* tile_order;      // tiles by descending cost (lb2_tile_order) or NULL
    lb2_conv_io io[2];
};

__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
// Arrive on a barrier of the leader CTA (address from mapa).  Default semantics (release at CTA scope) as in CUTLASS' ClusterBarrier:
// the .release.cluster / .acquire.cluster forms compile to MEMBAR.ALL.GPU + CCTL.IVALL around every arrive / poll (measured: the
// kernel ran 2x slower).  What the consumer (the pair's tensor cores, async proxy) needs is that the producer's cp.async writes have
// landed in its shared memory and are visible to the async proxy: cp.async.wait_group + fence.proxy.async by the writing thread
// before the arrive give exactly that; shared memory has no cache between the SM and the tensor-core read path.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) { mbar_wait(bar, parity); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {               // arrives on `bar` in BOTH CTAs of the pair once the prior MMAs are done
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// kind::f16 instruction descriptor for the pair: D = f32, A = B = f16, K-major, N >> 3 at [17,23), M = 256 -> 16 at [24,29)
__device__ __forceinline__ uint32_t make_idesc2(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24); }

struct Ring {                                            // position in a ring without integer division
    int s; uint32_t par; int n;
    __device__ __forceinline__ void next() { if (++s == n) { s = 0; par ^= 1u; } }
};

template <int NCOLS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) k_spconv_tc_pair(const Params p) {
    using C = Cfg<NCOLS>;
    constexpr int NB = C::NB, NACC = C::NACC;
    extern __shared__ unsigned char smem_raw[];
    const int M = p.d_mout ? min(*p.d_mout, p.mout_cap) : p.mout_cap;
    const int n_stiles = (M + 2 * BM - 1) / (2 * BM);           // super-tiles of 256 rows
    const int total = n_stiles * p.npass;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    // work item -> (tile, pass).  With a cost order (lb2_tile_order): item i = pass (i & 1) of the (i >> 1)-th most expensive tile, so the
    // item sequence is sorted by cost and the snake deal below is balanced to 1-3 %.  Without: pass-major, inside a pass the last
    // tiles of the mask-sorted row order (roughly the heaviest) first.
    const int pshift = (p.tile_order && p.npass == 2) ? 1 : 0;
    auto item_pass = [&](int item) { return p.tile_order ? (item & pshift) : (item >= n_stiles ? 1 : 0); };
    auto item_tile = [&](int item) {
        if (p.tile_order) return __ldg(p.tile_order + (item >> pshift));
        return n_stiles - 1 - (item >= n_stiles ? item - n_stiles : item);
    };
    // round j of the persistent loop in snake order (even rounds left to right, odd rounds right to left over the CTAs): with the
    // items sorted by cost every CTA alternates between a dearer and a cheaper item, so the per-CTA sums stay balanced (static LPT);
    // an item index >= total (last, partial round) is an empty tile for every role
    auto slot_item = [&](int jj) { return jj * npairs + ((jj & 1) ? npairs - 1 - pair : pair); };

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;               // same offset in both CTAs (same kernel, same dynamic size)
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t a_stage = 2u * A_TILE;                       // hi + lo image of 128 rows x 64 K-columns (= 2 A slots)
    const uint32_t b_base = base + (NA / 2) * a_stage;
    unsigned char* tail = gen + (size_t)(NA / 2) * a_stage + (size_t)NB * C::B_SLOT;
    float* slab = reinterpret_cast<float*>(tail);                                   // [8 warps][32][SLAB_PITCH]
    int* row_s = reinterpret_cast<int*>(tail + SLAB_BYTES);                          // [META][BM] own output rows
    uint32_t* mask_s = reinterpret_cast<uint32_t*>(row_s + META * BM);               // [META][BM] their neighbour bit masks
    uint32_t* wmask = mask_s + META * BM;                                            // [META][4] per-warp offset masks (union over the pair)
    uint64_t* bars = reinterpret_cast<uint64_t*>(wmask + 4 * META);
    uint32_t* misc = reinterpret_cast<uint32_t*>(bars + C::NBAR);
    float* aff_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(misc + 16) + 15) & ~uintptr_t(15));                              // [2][NCOLS] BN scale, shift
    const uint32_t bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * s; };                              // per CTA: its 128 producer threads (cp.async completion)
    auto full_ap = [&](int s) { return bar0 + 8u * (NA + s); };                      // leader's: the peer's relay
    auto empty_a = [&](int s) { return bar0 + 8u * (2 * NA + s); };                  // per CTA: multicast commit
    auto full_b = [&](int s) { return bar0 + 8u * (3 * NA + s); };                   // per CTA: bulk-copy transaction bytes
    auto full_bp = [&](int s) { return bar0 + 8u * (3 * NA + NB + s); };             // leader's: the peer's relay
    auto empty_b = [&](int s) { return bar0 + 8u * (3 * NA + 2 * NB + s); };         // per CTA: multicast commit
    auto acc_full = [&](int b) { return bar0 + 8u * (3 * NA + 3 * NB + b); };        // per CTA: multicast commit
    auto acc_empty = [&](int b) { return bar0 + 8u * (3 * NA + 3 * NB + NACC + b); };   // leader's: 16 drain warps of both CTAs
    auto meta_full = [&](int b) { return bar0 + 8u * (3 * NA + 3 * NB + 2 * NACC + b); };
    auto meta_empty = [&](int b) { return bar0 + 8u * (3 * NA + 3 * NB + 2 * NACC + META + b); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(full_a(s), 128); mbar_init(full_ap(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(full_b(s), 1); mbar_init(full_bp(s), 1); mbar_init(empty_b(s), 1); }
        for (int b = 0; b < NACC; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 16); }
        for (int b = 0; b < META; ++b) { mbar_init(meta_full(b), 1); mbar_init(meta_empty(b), 258); }   // warp 4 + loader + 256 drain threads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_affine(aff_s, p.scale, p.shift, NCOLS);
    if (warp == 4) {                                                                 // same warp in both CTAs, same destination offset
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc[0])), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();                                                              // barriers of both CTAs initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_d = misc[0];
    const float out_scale = __ldg(reinterpret_cast<const float*>(p.wpacked) + 1);
    auto tile_kmask = [&](int b) { return wmask[b * 4] | wmask[b * 4 + 1] | wmask[b * 4 + 2] | wmask[b * 4 + 3]; };

    if (warp < 4) {
        // =========================== WG0: A producers ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int t = threadIdx.x;
        const int sub = t & 3, rbase = t >> 2;                          // 16-byte chunk inside the half row / first of this thread's 4 rows
        int j = 0;
        Ring ri{0, 0u, NA};
        auto fetch_row = [&](int item, uint32_t r) {                    // output row of slot t of CTA r's tile in work item `item`
            if (item >= total) return -1;
            const int slot = item_tile(item) * 2 * BM + (int)r * BM + t;
            return (slot < M) ? (p.row_perm ? __ldg(p.row_perm + slot) : slot) : -1;
        };
        auto fetch_mask = [&](int row) -> uint32_t {                   // offsets a row has a neighbour at
            if (row < 0) return 0u;
            return p.row_mask ? __ldg(p.row_mask + row) : ((1u << p.kvol) - 1u);
        };
        int next_row = fetch_row(slot_item(0), rank), next_row_p = fetch_row(slot_item(0), rank ^ 1u);
        int next2_row = fetch_row(slot_item(1), rank), next2_row_p = fetch_row(slot_item(1), rank ^ 1u);
        uint32_t next_mask = fetch_mask(next_row), next_mask_p = fetch_mask(next_row_p);
        for (; j * npairs < total; ++j) {
            const int item = slot_item(j);
            const int b = j % META;
            const int pass = item_pass(item);
            const lb2_conv_io io = p.io[pass];
            if (j >= META) mbar_wait(meta_empty(b), ((j / META) - 1) & 1);
            {
                const int row = next_row;
                const uint32_t own = next_mask, both = next_mask | next_mask_p;
                next_row = next2_row; next_row_p = next2_row_p;
                next2_row = fetch_row(slot_item(j + 2), rank);         // prefetch two tiles ahead (rows), one tile ahead (masks)
                next2_row_p = fetch_row(slot_item(j + 2), rank ^ 1u);
                next_mask = fetch_mask(next_row); next_mask_p = fetch_mask(next_row_p);
                row_s[b * BM + t] = row;
                mask_s[b * BM + t] = own;
                const uint32_t wm = __reduce_or_sync(0xffffffffu, both);   // union over both tiles of the pair: identical in both CTAs
                if (lane == 0) wmask[b * 4 + warp] = wm;
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");
            if (t == 0) mbar_arrive(meta_full(b));
            const uint32_t kmask = tile_kmask(b);
            int myrows[4];
            uint32_t mymasks[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { myrows[q] = row_s[b * BM + rbase + 32 * q]; mymasks[q] = mask_s[b * BM + rbase + 32 * q]; }
            auto load_src = [&](int k, int (&dst)[4]) {              // neighbour rows of this thread's 4 tile rows at offset k
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    dst[q] = -1;
                    if ((mymasks[q] >> k) & 1u) dst[q] = p.nbr ? __ldg(p.nbr + (long long)k * p.nbr_stride + myrows[q]) : myrows[q];
                }
            };
            int src[4], nxt[4];
            uint32_t km = kmask;
            if (km) load_src(__ffs(km) - 1, src);
            while (km) {
                km &= km - 1;
                if (km) load_src(__ffs(km) - 1, nxt);                 // prefetch the next offset's rows behind this offset's copies
                for (int c2 = 0; c2 < 2 * p.nchunks; ++c2, ri.next()) {            // c2 = 2 * chunk + half
                    const int s = ri.s;
                    mbar_wait(empty_a(s), ri.par ^ 1u);
                    const uint32_t a_hi_u = base + (uint32_t)(s >> 1) * a_stage;
                    const int half = s & 1;                             // slots alternate halves: slot parity == c2 parity (NA even)
                    const int ch = c2 * 32 + sub * 8;                   // first of this thread's 8 input channels
                    const bool first = ch < p.c1;
                    const int cw = first ? p.c1 : p.c2;
                    const int co = first ? ch : ch - p.c1;
                    const __half* src_h = reinterpret_cast<const __half*>(first ? io.in1_h : io.in2_h);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = sw128(rbase + 32 * q, half * 4 + sub);
                        const bool ok = src[q] >= 0;
                        const __half* rp = src_h + (ok ? ((long long)src[q] * 2 * cw + co) : 0);
                        cp_async16(a_hi_u + off, rp, ok ? 16u : 0u);
                        cp_async16(a_hi_u + A_TILE + off, rp + (ok ? cw : 0), ok ? 16u : 0u);
                    }
                    cp_async_arrive_on(full_a(s));                      // published by the hardware when this thread's copies have landed
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) src[q] = nxt[q];
            }
        }
    } else if (warp < 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 4 && rank == 0) {
            // =========================== MMA issuer of the pair (whole warp converged; one elected lane issues) ===========================
            const uint32_t idesc = make_idesc2(NCOLS);
            int gcount = 0, j = 0;
            Ring rq{0, 0u, NA}, rb{0, 0u, NB};
            for (; j * npairs < total; ++j) {
            const int item = slot_item(j);
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                const int n_off = __popc(kmask);
                int in_group = 0, off_idx = 0;
                for (uint32_t km = kmask; km; km &= km - 1, ++off_idx) {
                    const int buf = gcount & (NACC - 1);
                    const uint32_t tmem_acc = tmem_d + (uint32_t)(buf * NCOLS);
                    if (in_group == 0 && gcount >= NACC) {
                        mbar_wait_cluster(acc_empty(buf), ((gcount / NACC) - 1) & 1);
                        tc_fence_after();
                    }
                    for (int c = 0; c < p.nchunks; ++c, rb.next()) {
                        const int sb = rb.s;
                        mbar_wait(full_b(sb), rb.par);                   // my half of the weight slot
                        mbar_wait_cluster(full_bp(sb), rb.par);          // the peer's half
                        const uint32_t b_hi = b_base + (uint32_t)sb * C::B_SLOT, b_lo = b_hi + C::B_HALF;
                        const uint64_t dbh0 = make_desc(b_hi), dbl0 = make_desc(b_lo);
#pragma unroll
                        for (int half = 0; half < 2; ++half, rq.next()) {
                            const int sa = rq.s;
                            mbar_wait(full_a(sa), rq.par);                  // my gathered rows
                            mbar_wait_cluster(full_ap(sa), rq.par);         // the peer's
                            tc_fence_after();
                            const uint32_t a_hi = base + (uint32_t)(sa >> 1) * a_stage, a_lo = a_hi + A_TILE;
                            const uint64_t dah0 = make_desc(a_hi) + 4u * (uint32_t)(sa & 1), dal0 = make_desc(a_lo) + 4u * (uint32_t)(sa & 1);
                            if (elect_one()) {
#pragma unroll
                                for (int k2 = 0; k2 < 2; ++k2) {          // +32 bytes per K step = +2 in the descriptor's address field
                                    const uint32_t kb = 2u * (uint32_t)(half * 2 + k2);
                                    const uint64_t dah = dah0 + 2u * k2, dal = dal0 + 2u * k2, dbh = dbh0 + kb, dbl = dbl0 + kb;
                                    umma2(tmem_acc, dah, dbh, idesc, (in_group | c | half | k2) ? 1u : 0u);
                                    umma2(tmem_acc, dal, dbh, idesc, 1);
                                    umma2(tmem_acc, dah, dbl, idesc, 1);
                                }
                                umma2_commit_both(empty_a(sa));
                            }
                            __syncwarp();
                        }
                        if (elect_one()) umma2_commit_both(empty_b(sb));
                        __syncwarp();
                    }
                    if (++in_group == p.group || off_idx == n_off - 1) {
                        if (elect_one()) umma2_commit_both(acc_full(buf));
                        __syncwarp();
                        in_group = 0;
                        ++gcount;
                    }
                }
                if (lane == 0) mbar_arrive(meta_empty(b));
                __syncwarp();
            }
        } else if (warp == 4 && lane == 0) {
            // =========================== rank 1: relay of "my weight half / my gathered rows have landed" to the leader,
            //                             in the order the MMA warp consumes them ===========================
            int j = 0;
            Ring r{0, 0u, NB}, ra{0, 0u, NA};
            const uint32_t leader_full_bp0 = map_to_cta(full_bp(0), 0), leader_full_ap0 = map_to_cta(full_ap(0), 0);
            for (; j * npairs < total; ++j) {
            const int item = slot_item(j);
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                for (uint32_t km = kmask; km; km &= km - 1) {
                    for (int c = 0; c < p.nchunks; ++c, r.next()) {
                        mbar_wait(full_b(r.s), r.par);
                        mbar_arrive_cluster(leader_full_bp0 + 8u * r.s);
#pragma unroll
                        for (int half = 0; half < 2; ++half, ra.next()) {
                            mbar_wait(full_a(ra.s), ra.par);
                            mbar_arrive_cluster(leader_full_ap0 + 8u * ra.s);
                        }
                    }
                }
                mbar_arrive(meta_empty(b));
            }
        } else if (warp == 5 && lane == 0) {
            // =========================== weight loader: this CTA's half of the columns ===========================
            int j = 0;
            Ring r{0, 0u, NB};
            const uint32_t self_full_bp0 = full_bp(0);
            for (; j * npairs < total; ++j) {
            const int item = slot_item(j);
                const int b = j % META;
                mbar_wait(meta_full(b), (j / META) & 1);
                const uint32_t kmask = tile_kmask(b);
                for (uint32_t km = kmask; km; km &= km - 1) {
                    const int k = __ffs(km) - 1;
                    for (int c = 0; c < p.nchunks; ++c, r.next()) {
                        const int s = r.s;
                        mbar_wait(empty_b(s), r.par ^ 1u);
                        const uint32_t dst = b_base + (uint32_t)s * C::B_SLOT;
                        const unsigned char* src = p.wpacked + PACK_HEADER + ((size_t)k * p.nchunks + c) * (4u * C::B_HALF) + (size_t)rank * C::B_HALF;
                        mbar_expect_tx(full_b(s), C::B_SLOT);
                        bulk_g2s(dst, src, C::B_HALF, full_b(s));                                  // hi tile, my rows
                        bulk_g2s(dst + C::B_HALF, src + 2u * C::B_HALF, C::B_HALF, full_b(s));     // lo tile, my rows
                        if (rank == 0) { /* the leader's own half needs no relay: full_bp is the peer's */ }
                    }
                }
                mbar_arrive(meta_empty(b));
            }
            (void)self_full_bp0;
        }
        __syncwarp();
    } else {
        // =========================== WG2 / WG3: drain (register-resident fp32 total) + epilogue of this CTA's 128 rows ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        constexpr int TOT = NCOLS / 2;                             // channels per drain warpgroup
        const int q4 = warp & 3;                                   // TMEM lane quarter
        const int cb = (warp >= 12) ? TOT : 0;                     // this warpgroup's first output channel
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        float* myslab = slab + (size_t)(warp - 8) * 32 * SLAB_PITCH;
        const uint32_t leader_acc_empty0 = map_to_cta(acc_empty(0), 0);
        float tot[TOT];
        int gcount = 0, j = 0;
        for (; j * npairs < total; ++j) {
            const int item = slot_item(j);
            const int b = j % META;
            const int pass = item_pass(item);
            const lb2_conv_io io = p.io[pass];
            mbar_wait(meta_full(b), (j / META) & 1);
            const uint32_t kmask = tile_kmask(b);
            const int n_off = __popc(kmask);
            const int n_groups = (n_off + p.group - 1) / p.group;
            const int* rows = row_s + b * BM + q4 * 32;
            int orows[4], gidx[4];                                  // the 4 rows this lane serves in the epilogue
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
                    prefetch_row_f32(io.residual, orows[i], NCOLS, cb, TOT, lane & 3);
                    if (!io.residual) prefetch_row_split(io.residual_h, orows[i], NCOLS, cb, TOT, lane & 3);
                    prefetch_row_f32(io.pre_add, orows[i], NCOLS, cb, TOT, lane & 3);
                    if (io.gate_table && io.gate_idx) prefetch_row_f32(io.gate_table, gidx[i], NCOLS, cb, TOT, lane & 3);
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
                drain_acc<TOT>(tmem_d + lane_base + (uint32_t)(buf * NCOLS + cb), tot, g == 0);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(leader_acc_empty0 + 8u * buf);   // accumulator free again: the MMA warp runs on while we finish
                ++gcount;
            }
            // ---- epilogue from registers, 16 channels at a time through the warp's slab (tc_common.cuh: epilogue_slabs) ----
            epilogue_slabs<TOT, 2>(tot, myslab, lane, orows, gidx, cb, NCOLS, out_scale, fl, io, aff_s);
            mbar_arrive(meta_empty(b));
        }
    }
    tc_fence_before();
    cluster_sync_all();                                                              // both CTAs are done with TMEM and with each other's barriers
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

}  // namespace tc5

bool lb2_spconv_tc5_supported(const lb2_conv_desc* d) {
    if (d->cout != 256 && d->cout != 128) return false;
    if ((long long)d->mout_cap * 2 * d->cout >= (1LL << 32)) return false;   // the epilogue indexes rows with 32-bit element offsets
    if (d->nbr != nullptr && d->row_mask == nullptr) return false;           // the pair's offset union is built from the callers' row masks
    if (d->kvol > 27) return false;
    for (int p = 0; p < d->npass; ++p) {
        if (!d->io[p].in1_h) return false;
        if (d->c2 > 0 && !d->io[p].in2_h) return false;
    }
    if ((d->c1 + d->c2) % tc::KC != 0 || d->c1 % 32 != 0) return false;      // whole 64-column chunks; a half never straddles in1/in2
    return true;
}

template <int NCOLS>
static int launch_pair(Lb2Handle* h, cudaStream_t s, const tc5::Params& p, int mout_cap, int npass) {
    const size_t smem = tc5::Cfg<NCOLS>::SMEM;
    static_assert(tc5::Cfg<NCOLS>::SMEM <= 227 * 1024, "shared memory budget");
    {
        cudaError_t e = lb2_configure_smem(h, NCOLS == 256 ? LB2_K_TC5_256 : LB2_K_TC5_128, tc5::k_spconv_tc_pair<NCOLS>, (int)(227 * 1024));
        if (e != cudaSuccess) return lb2_fail(h, LB2_ERR_CUDA, "k_spconv_tc_pair smem attribute: %s", cudaGetErrorString(e));
    }
    const long long stiles_cap = (long long)cdiv(mout_cap, 2 * tc::BM) * npass;
    const unsigned pairs = (unsigned)std::max<long long>(1, std::min<long long>(h->num_sms / 2, stiles_cap));
    tc5::k_spconv_tc_pair<NCOLS><<<2 * pairs, tc5::THREADS, smem, s>>>(p);
    LB2_POST_LAUNCH(h, "k_spconv_tc_pair");
    return LB2_OK;
}

int lb2_spconv_tc5_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget) {
    tc5::Params p;
    p.c1 = d->c1; p.c2 = d->c2; p.cout = d->cout; p.kvol = d->kvol; p.npass = d->npass;
    p.wpacked = (const unsigned char*)d->weight_packed;
    p.scale = d->scale; p.shift = d->shift; p.relu = d->relu;
    p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.d_mout = d->d_mout; p.mout_cap = d->mout_cap; p.row_perm = d->row_perm; p.row_mask = d->row_mask;
    p.nchunks = (d->c1 + d->c2) / tc::KC;
    const int steps_per_offset = 3 * ((d->c1 + d->c2 + 15) / 16);
    p.group = std::max(1, step_budget / steps_per_offset);
    p.io[0] = d->io[0]; p.io[1] = d->io[d->npass > 1 ? 1 : 0];
    p.tile_order = d->tile_order256;
    return d->cout == 256 ? launch_pair<256>(h, s, p, d->mout_cap, d->npass) : launch_pair<128>(h, s, p, d->mout_cap, d->npass);
}
