// K4 (variant B, persistent, 256 output channels) — specialisation of spconv_tc2.cu for Cout = 256, the layers that hold most
// of the FLOPs (L3 decoder, L4).  With N = 256 the accumulator alone fills half of TMEM, so the generic kernel has to
// keep the fp32 running total of the two-level accumulation in the other half and can neither ping-pong accumulators
// nor overlap drains/epilogue with the next MMAs (measured: ~25 % of these layers' time).  Here the running total lives
// in the REGISTERS of eight drain warps instead (setmaxnreg: producer / MMA warpgroups shrink to 56 registers, the two
// drain warpgroups grow to 200), TMEM holds two 256-column accumulators in ping-pong, and a group's drain as well as the
// whole epilogue overlap with the tensor-core work of the next group / tile.
// The gathered-A ring and the weight ring are separate: weights arrive in 2 slots of 64 K-columns (64 KB each, one bulk copy),
// gathered rows in 4 slots of 32 K-columns (the two halves of a 128-byte swizzled row image are disjoint sets of 16-byte
// chunks, so each half is its own pipeline slot), which gives the gathers a lookahead of two slots with one slot of slack.
//   WG0 warps 0-3   A producers (cp.async from the fp16 split companions only; neighbour rows re-read per offset)
//   WG1 warp 4 MMA issuer, warp 5 weight loader (warps 6,7 idle)
//   WG2 warps 8-11  drain + epilogue of output channels   0..127        WG3 warps 12-15: channels 128..255
// Same math and results as the per-tile kernel k_spconv_tc (tests compare them bit for bit).
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include "tc_common.cuh"

namespace tc3 {
using namespace tc;

constexpr int THREADS = 512;
constexpr int NCOLS = 256;
constexpr int MAX_KVOL = 27;
constexpr int NA = 4;                                 // A slots (half stages, 32 K-columns each)
constexpr int NB = 2;                                 // B slots (64 K-columns each)
constexpr int SLAB_PITCH = tc::EPI_PITCH;             // floats per slab row
constexpr int SLAB_BYTES = 8 * 32 * SLAB_PITCH * 4;   // 8 drain warps x 32 rows
constexpr int META = 4;                               // ring of per-tile metadata (row ids, offset masks).  Must exceed the cp.async
                                                      // lookahead A_LAG: a tile's last full_a arrival is issued up to A_LAG
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
    int nchunks, group, npass;
    lb2_conv_io io[2];
};

__global__ void __launch_bounds__(THREADS, 1) k_spconv_tc_n256(const Params p) {
    extern __shared__ unsigned char smem_raw[];
    const int M = p.d_mout ? min(*p.d_mout, p.mout_cap) : p.mout_cap;
    const int n_tiles = (M + BM - 1) / BM;
    const int total = n_tiles * p.npass;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item -> (tile, pass): pass-major (the two guidance passes read different feature tensors: one pass at a time keeps the
    // gathered working set inside the 126 MB L2), inside a pass the heaviest tiles first (the row order sorts rows by neighbour
    // mask, light to heavy), so the last, partially filled round of the persistent loop holds the cheapest tiles
    auto item_pass = [&](int item) { return item >= n_tiles ? 1 : 0; };
    auto item_tile = [&](int item) { return n_tiles - 1 - (item >= n_tiles ? item - n_tiles : item); };
    // round j of the persistent loop in snake order (even rounds left to right, odd rounds right to left over the CTAs): with the
    // items sorted by cost every CTA alternates between a dearer and a cheaper item, so the per-CTA sums stay balanced (static LPT);
    // an item index >= total (last, partial round) is an empty tile for every role
    auto slot_item = [&](int jj) { return jj * (int)gridDim.x + ((jj & 1) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x); };

    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* gen = smem_raw + (base - raw);
    const uint32_t b_tile = (uint32_t)NCOLS * 128u;
    const uint32_t a_stage = 2u * A_TILE;                       // hi + lo image of 128 rows x 64 K-columns (= 2 A slots)
    const uint32_t b_base = base + (NA / 2) * a_stage;          // weight slots follow the A images
    unsigned char* tail = gen + (size_t)(NA / 2) * a_stage + (size_t)NB * 2u * b_tile;
    float* slab = reinterpret_cast<float*>(tail);                                   // [8 warps][32][SLAB_PITCH]
    int* row_s = reinterpret_cast<int*>(tail + SLAB_BYTES);                          // [META][BM]
    uint32_t* mask_s = reinterpret_cast<uint32_t*>(row_s + META * BM);               // [META][BM] neighbour bit mask of each tile row
    uint32_t* wmask = mask_s + META * BM;                                            // [META][4] per-warp offset masks
    uint64_t* bars = reinterpret_cast<uint64_t*>(wmask + 4 * META);
    constexpr int NBAR = 2 * NA + 2 * NB + 4 + 2 * META;
    uint32_t* misc = reinterpret_cast<uint32_t*>(bars + NBAR);
    float* aff_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(misc + 16) + 15) & ~uintptr_t(15));                              // [2][NCOLS] BN scale, shift
    const uint32_t bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto empty_a = [&](int s) { return bar0 + 8u * (NA + s); };
    auto full_b = [&](int s) { return bar0 + 8u * (2 * NA + s); };
    auto empty_b = [&](int s) { return bar0 + 8u * (2 * NA + NB + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + 2 + b); };
    auto meta_full = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + 4 + b); };
    auto meta_empty = [&](int b) { return bar0 + 8u * (2 * NA + 2 * NB + 4 + META + b); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { mbar_init(full_a(s), 128); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 256); }        // 8 drain warps
        for (int b = 0; b < META; ++b) { mbar_init(meta_full(b), 1); mbar_init(meta_empty(b), 258); }   // MMA + loader + 256 drain threads
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    stage_affine(aff_s, p.scale, p.shift, NCOLS);
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&misc[0])), "r"(512u) : "memory");
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
        // =========================== WG0: A producers ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        const int t = threadIdx.x;
        const int sub = t & 3, rbase = t >> 2;                          // 16-byte chunk inside the half row / first of this thread's 4 rows
        int j = 0;
        Ring ri{0, 0u, NA};
        auto fetch_row = [&](int item) {
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
            const int pass = item_pass(item);
            const lb2_conv_io io = p.io[pass];
            if (j >= META) mbar_wait(meta_empty(b), ((j / META) - 1) & 1);
            {
                const int row = next_row;
                const uint32_t have = next_mask;
                next_row = next2_row;
                next2_row = fetch_row(slot_item(j + 2));            // prefetch two tiles ahead (row), one tile ahead (its mask)
                next_mask = fetch_mask(next_row);
                row_s[b * BM + t] = row;
                uint32_t found = have;
                if (!p.row_mask) {                                      // no mask from the caller: find the populated offsets by reading the map
                    found = 0;
                    for (int k0 = 0; k0 < p.kvol; k0 += 9) {
                        int v[9];
#pragma unroll
                        for (int q = 0; q < 9; ++q) {
                            const int k = k0 + q;
                            v[q] = -1;
                            if (k < p.kvol && row >= 0) v[q] = p.nbr ? __ldg(p.nbr + (long long)k * p.nbr_stride + row) : row;
                        }
#pragma unroll
                        for (int q = 0; q < 9; ++q)
                            if (k0 + q < p.kvol && v[q] >= 0) found |= 1u << (k0 + q);
                    }
                }
                mask_s[b * BM + t] = found;
                const uint32_t wm = __reduce_or_sync(0xffffffffu, found);
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
        if (warp == 4) {
            // =========================== MMA issuer (whole warp converged; one elected lane issues) ===========================
            const uint32_t idesc = make_idesc(NCOLS);
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
                    const int buf = gcount & 1;
                    const uint32_t tmem_acc = tmem_d + (uint32_t)(buf * NCOLS);
                    if (in_group == 0 && gcount >= 2) {
                        mbar_wait(acc_empty(buf), ((gcount >> 1) - 1) & 1);
                        tc_fence_after();
                    }
                    for (int c = 0; c < p.nchunks; ++c, rb.next()) {
                        const int sb = rb.s;
                        mbar_wait(full_b(sb), rb.par);
                        const uint32_t b_hi = b_base + (uint32_t)sb * 2u * b_tile, b_lo = b_hi + b_tile;
                        const uint64_t dbh0 = make_desc(b_hi), dbl0 = make_desc(b_lo);
#pragma unroll
                        for (int half = 0; half < 2; ++half, rq.next()) {
                            const int sa = rq.s;
                            mbar_wait(full_a(sa), rq.par);
                            tc_fence_after();
                            const uint32_t a_hi = base + (uint32_t)(sa >> 1) * a_stage, a_lo = a_hi + A_TILE;
                            const uint64_t dah0 = make_desc(a_hi), dal0 = make_desc(a_lo);
                            if (elect_one()) {
#pragma unroll
                                for (int k2 = 0; k2 < 2; ++k2) {          // +32 bytes per K step = +2 in the address field
                                    const uint32_t ks = (uint32_t)(half * 2 + k2);
                                    const uint64_t dah = dah0 + 2u * ks, dal = dal0 + 2u * ks, dbh = dbh0 + 2u * ks, dbl = dbl0 + 2u * ks;
                                    umma(tmem_acc, dah, dbh, idesc, (in_group | c | (int)ks) ? 1u : 0u);
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
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int q4 = warp & 3;                                   // TMEM lane quarter
        const int cb = (warp >= 12) ? 128 : 0;                     // this warpgroup's first output channel
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        float* myslab = slab + (size_t)(warp - 8) * 32 * SLAB_PITCH;
        float tot[128];
        int gcount = 0, j = 0;
        for (; j * (int)gridDim.x < total; ++j) {
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
            if (n_groups == 0) {
#pragma unroll
                for (int q = 0; q < 128; ++q) tot[q] = 0.f;
            }
            for (int g = 0; g < n_groups; ++g) {
                const int buf = gcount & 1;
                mbar_wait(acc_full(buf), (gcount >> 1) & 1);
                tc_fence_after();
                drain_acc<128>(tmem_d + lane_base + (uint32_t)(buf * NCOLS + cb), tot, g == 0);
                tc_fence_before();
                mbar_arrive(acc_empty(buf));                       // accumulator free again: the MMA warp runs on while we finish
                ++gcount;
            }
            // ---- epilogue from registers, 16 channels at a time through the warp's slab (tc_common.cuh: epilogue_slabs) ----
            epilogue_slabs<128, 2>(tot, myslab, lane, orows, gidx, cb, NCOLS, out_scale, epi_flags(io, p.relu), io, aff_s);
            mbar_arrive(meta_empty(b));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512u) : "memory");
}

static size_t smem_bytes() {
    return 1024 + (size_t)(NA / 2) * 2 * A_TILE + (size_t)NB * 2 * NCOLS * 128 + SLAB_BYTES + META * BM * sizeof(int) +
           META * BM * sizeof(uint32_t) + 4 * META * sizeof(uint32_t) + (2 * NA + 2 * NB + 4 + 2 * META) * 8 + 64 + 16 + 2 * NCOLS * sizeof(float);
}

}  // namespace tc3

bool lb2_spconv_tc3_supported(const lb2_conv_desc* d) {
    if (d->cout != 256) return false;
    if ((long long)d->mout_cap * 2 * d->cout >= (1LL << 32)) return false;   // the epilogue indexes rows with 32-bit element offsets
    for (int p = 0; p < d->npass; ++p) {
        if (!d->io[p].in1_h) return false;
        if (d->c2 > 0 && !d->io[p].in2_h) return false;
    }
    if ((d->c1 + d->c2) % tc::KC != 0 || d->c1 % 32 != 0) return false;      // whole 64-column chunks; a half never straddles in1/in2
    return tc3::smem_bytes() <= 227 * 1024;
}

int lb2_spconv_tc3_launch(Lb2Handle* h, cudaStream_t s, const lb2_conv_desc* d, int step_budget) {
    tc3::Params p;
    p.c1 = d->c1; p.c2 = d->c2; p.cout = d->cout; p.kvol = d->kvol; p.npass = d->npass;
    p.wpacked = (const unsigned char*)d->weight_packed;
    p.scale = d->scale; p.shift = d->shift; p.relu = d->relu;
    p.nbr = d->nbr; p.nbr_stride = d->nbr_stride; p.d_mout = d->d_mout; p.mout_cap = d->mout_cap; p.row_perm = d->row_perm; p.row_mask = d->row_mask;
    p.nchunks = (d->c1 + d->c2 + tc::KC - 1) / tc::KC;
    const int steps_per_offset = 3 * ((d->c1 + d->c2 + 15) / 16);
    p.group = std::max(1, step_budget / steps_per_offset);
    p.io[0] = d->io[0]; p.io[1] = d->io[d->npass > 1 ? 1 : 0];
    const size_t smem = tc3::smem_bytes();
    {
        cudaError_t e = lb2_configure_smem(h, LB2_K_TC3, tc3::k_spconv_tc_n256, (int)(227 * 1024));
        if (e != cudaSuccess) return lb2_fail(h, LB2_ERR_CUDA, "k_spconv_tc_n256 smem attribute: %s", cudaGetErrorString(e));
    }
    const long long tiles_cap = (long long)cdiv(d->mout_cap, tc::BM) * d->npass;
    const unsigned grid = (unsigned)std::min<long long>(h->num_sms, tiles_cap);
    tc3::k_spconv_tc_n256<<<grid, tc3::THREADS, smem, s>>>(p);
    LB2_POST_LAUNCH(h, "k_spconv_tc_n256");
    return LB2_OK;
}
