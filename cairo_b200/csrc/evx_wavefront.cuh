// evx_wavefront.cuh -- K3, the encoder's serial spine: intra search in the frame under
// construction (motion.cpp:354-419), classify (encode.cpp:17-67), encode_block
// (encode.cpp:69-163) and the reconstruction loop (decode.cpp:15-144).
//
// Macroblocks depend on each other in raster order through the intra search (it reads the
// neighbours' reconstructions, and stale samples of the ring slot -- SURVEY H3).  The legal
// schedule is a wavefront: (bx,by) after (bx-1,by) and (min(bx+2,W-1),by-1).
//
// Mapping: ONE CTA PER MACROBLOCK ROW, warp specialised --
//   warps 0..CW-1  compute (CW = 8 or 4, EvxK3Cfg): the eight outer candidates of a 3x3 search round, 8 / CW per warp (the
//               centre is the running best, whose cost is already known), then the transform
//               path of the macroblock;
//   warp  CW    publisher + block loader.  Each time a macroblock of the row completes it first
//               fences and releases progress[by] so the row below can follow, then stages the
//               macroblock after next: source block, K2's inter results and their predictions,
//               the stale columns of the row below -- nothing that depends on the row above;
//   warp  CW+1  column loader: slides the search window (a ring of 8 macroblock columns in
//               shared memory) -- the 16 new columns of the three rows above as soon as the row
//               above has published them.  A loader of its own, because this is the row-to-row
//               critical path: with one loader doing both, staging macroblock n+1 queued behind
//               the wait for column n+2 and every row trailed its predecessor by 4 macroblock
//               times instead of 2 + the hand-off;
// The left-neighbour dependency is thus inside the CTA (its reconstruction is written straight
// into the window), and the inter-row latency (flag + L2 round trip) is paid once per row
// instead of once per macroblock: frame time ~ (W + 3(H-1)) * T_mb + (H-1) * latency.
//
// THE FRAME PIPELINE (frames of a stream in flight on the device at once, evxgpu.cu submit_pipelined).  A frame is
// THREE launches that run side by side, each following the one before it through counters in device memory:
//   evx_search_follow    the frame's inter search.  Persistent CTAs take macroblock ROWS by ticket, in order; a row's
//                        (macroblock, reference) items are claimed in column order through a counter k2c[y] by the
//                        CTA's warps, each item once the previous frame is final around it (evx_gate_prev); every
//                        result is stamped;
//   evx_wavefront        the wavefront rows (this file's main kernel).  A row's block loader waits for the stamps of
//                        the macroblock it stages -- after it has claimed, and run itself, whatever item of ITS OWN row
//                        up to that macroblock nobody has taken (so the stamps it waits for are always in the hands of
//                        running warps, whether or not a search CTA is resident);
//   evx_deblock_follow   the deblocking filter behind the wavefront.  The sweep decomposes into independent 8x8 tiles
//                        (evx_kernels.cuh, K4); the tiles of macroblock (X, Y) -- luma crossings {2X, 2X+1} x {2Y, 2Y+1},
//                        chroma (X, Y), plus the frame's right / bottom border tiles in the last column / row -- touch
//                        macroblocks (X-1..X, Y-1..Y), whose unfiltered samples the intra search reads last from
//                        macroblock (X+2, Y+3).  A JOB is four tile columns of one tile row; persistent one-warp CTAs take
//                        TILE ROWS by ticket and run a row's jobs as the wavefront passes them (claimed through jc[Y],
//                        a compare-and-swap: a claimed job never waits), publishing dbk[Y] = base + filtered tile
//                        columns.  The wavefront kernel's last row takes whatever jobs are left when the rows are done.
// Nobody waits for work that could still need an SM slot the waiter holds: tickets are claimed in dependency order, the
// wavefront kernel has a fallback for everything it consumes, and the other waits point at the previous frame, whose
// kernels were launched -- and have started -- before this frame's.  Any number of streams and processes may share
// the device.  Cross-frame counters hold frame base + count and are compared cyclically; every wait is bounded.
//   * the NEXT frame reads this one as a reference around (bx, by): samples of macroblocks (bx-2..bx+2, by-2..by+2),
//     final once the tiles of columns <= bx+3 in tile rows by-2 .. by+3 are done.  Its search items (and, in an intra
//     frame, its block loaders) wait for dbk[by-2 .. by+3] >= min(bx+4, W), six lanes polling one row each.  That
//     one rule also covers the write-after-read side (a frame overwrites the ring slot of frame n-R, which frame
//     n-1 still searches) and the stale samples the intra search reads from that slot: every frame
//     trails its predecessor by the same ~30 wavefront steps at every macroblock, and so transitively all older ones.
#pragma once

#include "evx_kernels.cuh"

// Compute warps of a row CTA, per build of the kernel (the eight outer cells of a search round and the eight sub-pel
// directions are spread over them: 8 / CW each).  Eight warps make the shortest macroblock, so the kernel that has the
// device to itself uses eight (1.65 against 1.79 ms per 1080p frame).  Every warp repeats the selection and the state
// update of a round, so four warps execute a third fewer instructions per macroblock -- and next to other kernels
// instruction supply is the limit: 2 432 -> 2 844 frames/s in the eight-slot pipeline, 2 958 -> 3 256 with
// 16 streams.  So the kernel that shares the device uses four.
#ifndef EVX_K3_CW_ALONE
#define EVX_K3_CW_ALONE 8
#endif
#ifndef EVX_K3_CW_SHARED
#define EVX_K3_CW_SHARED 4
#endif
#ifndef EVX_K3_MINCTAS
#define EVX_K3_MINCTAS 2            // CTAs per SM the register budget of the sharing build is set for
#endif
template <int MINCTAS> struct EvxK3Cfg
{
    static constexpr int CW = MINCTAS == 1 ? EVX_K3_CW_ALONE : EVX_K3_CW_SHARED;      // compute warps
    static constexpr int CT = CW * 32;                                               // compute threads
    static constexpr int NT = CT + 64;                                               // + block loader + column loader
};
// The 384 elements of a macroblock over the compute threads.  UNROLLED (a constant in scope: the latency-optimised build of
// the kernel): a fully unrolled outer loop around an inner loop that runs at most once, so that with 256 threads the two
// passes (the second one only for threads < 128) sit in one basic block and their loads overlap.  Otherwise the plain
// strided loop: a third of the code.  See evx_k3_compute for when which form is used.  (CT: the compute threads, in scope.)
#define EVX_K3_FOR384(e) _Pragma("unroll") for (int e##_it = 0; e##_it < (UNROLLED ? (384 + CT - 1) / CT : 1); ++e##_it) \
                         for (int e = tid + e##_it * CT; e < 384; e += (UNROLLED ? 384 : CT))
#define EVX_K3_ROWS 80            // window rows py-48 .. py+31
#define EVX_K3_CROWS 40
#define EVX_MAXREF 7

struct EvxK3Smem
{
    uint32_t wy[EVX_K3_ROWS * EVX_RING_PWY];
    uint32_t wu[EVX_K3_CROWS * EVX_RING_PWC];
    uint32_t wv[EVX_K3_CROWS * EVX_RING_PWC];
    int16_t src[2][384];                      // block-major source macroblocks, double buffered
    int16_t ipred[2][EVX_MAXREF][384];        // predictions of K2's candidates, block-major
    int4 idesc[2][EVX_MAXREF];
    int isad[2][EVX_MAXREF + 1];
    EvxMbShared sh;
    int4 cand[2][16];                         // sub-pel tests: {sad, mad, 0, legal}
    int2 cand2[2][16];                        // full-pel rounds: raw {sad, mad} per cell
    uint64_t full[2], fullb[2], full2[2], empty[2];   // full: block data; fullb: window columns <= n+1; full2: column n+2
    uint64_t k2bar;                           // the block loader's own search window (it serves the search queue when it must)
    __align__(128) uint8_t k2win[EVX_K2W_BYTES];
    int last_motion, last_coded;              // K8 bookkeeping (thread 0)
};

// the CTA's current ticket
struct EvxFrameCtl { int row; };
#define EVX_FRAME_CTL_BYTES 128
#define EVX_FRAME_SMEM (EVX_FRAME_CTL_BYTES + sizeof(EvxK3Smem))

__device__ __forceinline__ void evx_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(evx_smem_addr(bar)) : "memory");
}

template <int CT> __device__ __forceinline__ void evx_compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }

// int16 sample of the ring window
__device__ __forceinline__ int evx_ring_y(const EvxK3Smem &S, int x, int wrow) { return reinterpret_cast<const int16_t *>(S.wy)[wrow * (EVX_RING_PWY * 2) + (x & 127)]; }
__device__ __forceinline__ int evx_ring_c(const uint32_t *pl, int cx, int wrow) { return reinterpret_cast<const int16_t *>(pl)[wrow * (EVX_RING_PWC * 2) + (cx & 63)]; }

// ---------------------------------------------------------------- the previous frame, and the search role

#ifndef EVX_GATE_POLL_NS
#define EVX_GATE_POLL_NS 200      // back-off of the poll of the previous frame's deblocking counters
#endif
// Waits until the previous frame of the stream is FINAL (reconstructed and deblocked) in every sample macroblock
// (bx, by) of this frame reads as a reference: six lanes poll tile rows by-2 .. by+3 (header comment).
__device__ __forceinline__ void evx_gate_prev(const EvxK3Params &p, int bx, int by, int lane)
{
    if (!p.prev_dbk) return;
    const unsigned int need = p.prev_base + (unsigned int) min(bx + 4, p.g.mbw);
    const int row = max(0, min(by + 3 - min(lane, 5), p.g.mbh - 1));
    unsigned int v;
    EVX_BOUNDED_WAIT(p.wait, (v = evx_ld_relaxed_u32(p.prev_dbk + row), __all_sync(0xFFFFFFFFu, (int) (v - need) >= 0)), EVX_GATE_POLL_NS, 4u, need, v, (unsigned int) row);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// How many tile columns the previous frame still lacks around macroblock (bx, row) (<= 0: final there): six lanes look at
// tile rows row-2 .. row+3 (evx_gate_prev).
__device__ __forceinline__ int evx_gate_shortfall(const EvxK3Params &p, int bx, int row, int lane)
{
    const unsigned int want = p.prev_base + (unsigned int) min(bx + 4, p.g.mbw);
    const int short_by = (int) (want - evx_ld_relaxed_u32(p.prev_dbk + max(0, min(row + 3 - min(lane, 5), p.g.mbh - 1))));
    return __reduce_max_sync(0xFFFFFFFFu, short_by);
}

// A call, not inlined: the tile's 64 samples live in registers, and inlined into the wavefront kernel they would raise the
// pressure of every other path of it (spills in the two-CTAs-per-SM budget); as a function the cost stays in here.
// (arguments by value: a reference to the caller's locals would put those on its stack)
__device__ __noinline__ void evx_deblock_tile_call(int16_t *y, int16_t *u, int16_t *v, int w, int h, const EvxDesc *table, int comp, int tx, int ty)
{
    EvxPlanes pl; pl.y = y; pl.u = u; pl.v = v;
    EvxGeom g; g.w = w; g.h = h; g.vw = w; g.vh = h; g.mbw = w >> 4; g.mbh = h >> 4;
    evx_deblock_tile(pl, g, table, comp, tx, ty);
}

// One deblocking job: tile columns [x0, cnt) of tile row Y, one tile per lane and pass; then the job's count is
// published in order (the job before it in the row was claimed earlier and is in a running warp's hands).
__device__ __noinline__ void evx_deblock_job(const EvxK3Params &p, int Y, int x0, int cnt, int lane)
{
    const EvxGeom g = p.g;
    if (p.deblocking)
    {
        const EvxPlanes cur = p.ring[(int) (p.frame_index % (uint32_t) p.R)];
        const int last = cnt == g.mbw ? 1 : 0, bot = Y == g.mbh - 1 ? 1 : 0;
        const int nlx = 2 * (cnt - x0) + last, ncx = (cnt - x0) + last;
        const int nl = nlx * (2 + bot), nc = ncx * (1 + bot);
        for (int k = lane; k < nl + 2 * nc; k += 32)
        {
            if (k < nl) evx_deblock_tile_call(cur.y, cur.u, cur.v, g.w, g.h, p.table, 0, 2 * x0 + k % nlx, 2 * Y + k / nlx);
            else { const int c = (k - nl) % nc; evx_deblock_tile_call(cur.y, cur.u, cur.v, g.w, g.h, p.table, 1 + (k - nl) / nc, x0 + c % ncx, Y + c / ncx); }
        }
    }
    __syncwarp();
    if (lane == 0)
    {
        if (x0 > 0) EVX_BOUNDED_WAIT(p.wait, (int) (evx_ld_relaxed_u32(p.dbk + Y) - (p.dbk_base + (unsigned int) x0)) >= 0, 200, 8u, (unsigned int) Y, (unsigned int) x0, 0u);
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.dbk + Y), "r"(p.dbk_base + (unsigned int) cnt) : "memory");
    }
    __syncwarp();
}

#ifndef EVX_DBK_CHUNK
#define EVX_DBK_CHUNK 4           // tile columns per deblocking job
#endif
#ifndef EVX_DBK_POLL_NS
#define EVX_DBK_POLL_NS 4000      // longest back-off of the deblocking follower's poll
#endif

// per-row counters in p.sync: progress[y] (macroblocks of wavefront row y complete), k2c[y] (search items of row y
// claimed), jc[Y] (deblocking jobs of tile row Y claimed)
__device__ __forceinline__ int *evx_progress(const EvxK3Params &p) { return p.sync + 2; }
__device__ __forceinline__ int *evx_k2c(const EvxK3Params &p) { return p.sync + 2 + p.g.mbh; }
__device__ __forceinline__ int *evx_jc(const EvxK3Params &p) { return p.sync + 2 + 2 * p.g.mbh; }

// Scans the tile rows (lane = row within a chunk of 32) for a deblocking job the wavefront has passed -- tile columns
// [4c, 4c+4) of tile row Y may be filtered once macroblock (min(4c+5, W-1), min(Y+3, H-1)) is complete -- claims one by
// compare-and-swap on jc[Y] (a claimed job must never wait for the wavefront) and runs it.  `rot` spreads the warps
// over the ready rows.  Returns 1 if a job was done, 0 if none is ready, -1 if every job of the frame is claimed.
// (The wavefront kernel's last row: whatever the deblocking follower has not taken.)
__device__ __forceinline__ int evx_serve_deblock(const EvxK3Params &p, int lane, unsigned int rot)
{
    const int H = p.g.mbh, W = p.g.mbw, nch = (W + EVX_DBK_CHUNK - 1) / EVX_DBK_CHUNK;
    int *jc = evx_jc(p);
    const int *progress = evx_progress(p);
    bool all = true;
    for (int base = 0; base < H; base += 32)
    {
        const int Y = base + lane;
        int c = nch;
        if (Y < H) c = evx_ld_relaxed(jc + Y);
        bool ready = false;
        if (c < nch) ready = evx_ld_relaxed(progress + min(Y + 3, H - 1)) >= min(min(c * EVX_DBK_CHUNK + EVX_DBK_CHUNK, W) + 1, W - 1) + 1;
        if (__any_sync(0xFFFFFFFFu, c < nch)) all = false;
        unsigned int mask = __ballot_sync(0xFFFFFFFFu, ready);
        while (mask)
        {
            const unsigned int r = rot & 31u, m2 = (mask >> r) | (r ? mask << (32u - r) : 0u);       // first ready lane at or after `rot`
            const int l = (int) ((__ffs(m2) - 1 + r) & 31u);
            int ok = 0;
            if (lane == l) ok = atomicCAS(jc + Y, c, c + 1) == c;
            ok = __shfl_sync(0xFFFFFFFFu, ok, l);
            if (ok)
            {
                const int Yj = __shfl_sync(0xFFFFFFFFu, Y, l), cj = __shfl_sync(0xFFFFFFFFu, c, l);
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                evx_deblock_job(p, Yj, cj * EVX_DBK_CHUNK, min(cj * EVX_DBK_CHUNK + EVX_DBK_CHUNK, W), lane);
                return 1;
            }
            mask &= ~(1u << l);
        }
    }
    return all ? -1 : 0;
}

// One search item of row y: item index it = x * nref + ref.
__device__ __forceinline__ void evx_run_search_item(const EvxK3Params &p, int y, int it, int lane, uint8_t *win, uint64_t *bar, uint32_t &phase)
{
    const int nref = p.R - 1, nmb = p.g.mbw * p.g.mbh, bx = it / nref, ref = it - bx * nref;
    const int rslot = (int) ((p.frame_index + (uint32_t) p.R - (uint32_t) (ref + 1)) % (uint32_t) p.R);       // common.cpp:192-195
    evx_k2_item<2>(&p.maps.m[ref * 3], p.src, p.ring[rslot], p.g, p.thr, bx, y, ref, lane, win, bar, phase,
                p.inter + (size_t) ref * nmb + (size_t) y * p.g.mbw + bx, p.counters, p.stamp);
}

// The block loader's part: every item of ITS row up to macroblock n must be claimed before it waits for their stamps.
__device__ __noinline__ void evx_claim_own_items(const EvxK3Params &p, int by, int n, int lane, uint8_t *win, uint64_t *bar, uint32_t &phase)
{
    const int nref = p.R - 1, until = (n + 1) * nref, per_row = p.g.mbw * nref;
    int *k2c = evx_k2c(p) + by;
    for (;;)
    {
        int it = per_row;
        if (lane == 0 && evx_ld_relaxed(k2c) < until) it = atomicAdd(k2c, 1);
        it = __shfl_sync(0xFFFFFFFFu, it, 0);
        if (it >= per_row) break;
        evx_gate_prev(p, it / nref, by, lane);
        evx_run_search_item(p, by, it, lane, win, bar, phase);
    }
}

// ---------------------------------------------------------------- block loader warp

__device__ __forceinline__ void evx_k3_block_loader(EvxK3Smem &S, const EvxK3Params &p, int by, int lane)
{
    const EvxGeom g = p.g;
    const int nmb = g.mbw * g.mbh, cw = g.w >> 1, py = by * EVX_MB;
    const int dest = (int) (p.frame_index % (uint32_t) p.R);
    const EvxPlanes cur = p.ring[dest];
    int *progress = p.sync + 2;
    const int nref = p.frame_type == 1 ? p.R - 1 : 0;
    // macroblock m of this row is complete (its `empty` phase has passed; the compute warps cannot
    // be more than one macroblock past the one being staged, so a phase is never lapped): make the
    // row's writes visible device-wide, then advance progress[by]
    auto publish = [&](int m)
    {
        evx_mbar_wait(&S.empty[m & 1], (uint32_t) ((m >> 1) & 1));
        if (lane == 0) evx_st_release(progress + by, m + 1);     // release is cumulative over what this thread observed through the barrier
    };
    uint32_t k2phase = 0;
#ifdef EVX_K3_LOADER_STATS
    long long ls_own = 0, ls_own_cyc = 0, ls_stamp_cyc = 0;
#endif
    bool stamps_seen = false;       // the stamps of the macroblock about to be staged were already seen (looked at one macroblock early)

    for (int n = 0; n < g.mbw; ++n)
    {
        const int slot = n & 1, px = n * EVX_MB, mb = by * g.mbw + n;
        if (n >= 2) publish(n - 2);          // also frees this slot's staging buffers
        // An intra frame has no search whose items wait for the previous frame, but it overwrites the ring slot frames
        // before it still read and reads that slot's stale samples: the same gate, taken here.
        if (!nref) evx_gate_prev(p, n, by, lane);

        // source macroblock -> block-major
        {
            int row = lane >> 1, half = lane & 1;
            uint4 v = __ldcg(reinterpret_cast<const uint4 *>(p.src.y + (size_t) (py + row) * g.w + px + 8 * half));
            *reinterpret_cast<uint4 *>(&S.src[slot][((row >> 3) * 2 + half) * 64 + (row & 7) * 8]) = v;
            if (lane < 16)
            {
                const int16_t *pl = lane < 8 ? p.src.u : p.src.v;
                uint4 c = __ldcg(reinterpret_cast<const uint4 *>(pl + (size_t) ((py >> 1) + (lane & 7)) * cw + (px >> 1)));
                *reinterpret_cast<uint4 *>(&S.src[slot][256 + (lane >> 3) * 64 + (lane & 7) * 8]) = c;
            }
        }
        // K2's candidates and their predictions (encode.cpp:110-141), so that classification
        // never waits on global memory
        if (nref && p.fuse_k2)
        {   // the search results of this macroblock: every item up to its own must be claimed (this warp claims and runs what
            // nobody has yet), then one lane per reference waits for the stamps (a claimed item is in a running warp's hands)
            if (!stamps_seen)
            {
#ifdef EVX_K3_LOADER_STATS
                const long long t0 = clock64();
                const int c0 = evx_ld_relaxed(evx_k2c(p) + by);
#endif
                evx_claim_own_items(p, by, n, lane, S.k2win, &S.k2bar, k2phase);
#ifdef EVX_K3_LOADER_STATS
                const long long t1 = clock64();
                if (evx_ld_relaxed(evx_k2c(p) + by) != c0 && t1 - t0 > 2000) { ls_own++; ls_own_cyc += t1 - t0; }
#endif
                const uint32_t *st = &p.inter[(size_t) min(lane, nref - 1) * nmb + mb].stamp;
                EVX_BOUNDED_WAIT(p.wait, __all_sync(0xFFFFFFFFu, evx_ld_relaxed_u32(st) == p.stamp), 100, 5u, (unsigned int) mb, p.stamp, 0u);
#ifdef EVX_K3_LOADER_STATS
                ls_stamp_cyc += clock64() - t1;
#endif
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        for (int r = 0; r < nref; ++r)
        {
            const EvxInterResult *ir = p.inter + (size_t) r * nmb + mb;
            int4 raw = __ldcg(reinterpret_cast<const int4 *>(&ir->desc));
            int isad = __ldcg(&ir->sad);
            if (lane == 0) { S.idesc[slot][r] = raw; S.isad[slot][r] = isad; }
            EvxDesc d; d.w0 = raw.x; d.w1 = raw.y; d.w2 = raw.z; d.w3 = raw.w;
            const int type = d.type();
            const int rslot = (int) ((p.frame_index + (uint32_t) p.R - (uint32_t) (r + 1)) % (uint32_t) p.R);
            const EvxPlanes ref = p.ring[rslot];
            const int mx = (type & EVX_T_MOTION) ? d.mx() : 0, my = (type & EVX_T_MOTION) ? d.my() : 0;
            const bool sp = (type & EVX_T_MOTION) && d.sp_pred();
            int dx = 0, dy = 0;
            if (sp) evx_frac_direction(d.sp_index(), dx, dy);
            const int bxp = px + mx, byp = py + my;
            if (!sp && (bxp & 15) == 0)
            {   // (zero motion above all: 128-bit loads, two chunks per lane; warp-uniform)
                evx_pred_chunk16(S.ipred[slot][r], ref, g, bxp, byp, lane);
                if (lane < 16) evx_pred_chunk16(S.ipred[slot][r], ref, g, bxp, byp, 32 + lane);
                continue;
            }
            int a[12], b[12];
#pragma unroll
            for (int k = 0; k < 12; ++k)
            {
                int comp, x, y;
                evx_mb_pos(lane + 32 * k, comp, x, y);
                b[k] = 0;
                if (comp == 0)
                {
                    a[k] = __ldcg(ref.y + (size_t) (byp + y) * g.w + bxp + x);
                    if (sp) b[k] = __ldcg(ref.y + (size_t) (byp + dy + y) * g.w + bxp + dx + x);
                }
                else
                {
                    const int16_t *pl = comp == 1 ? ref.u : ref.v;
                    a[k] = __ldcg(pl + (size_t) ((byp >> 1) + y) * cw + (bxp >> 1) + x);
                    if (sp) b[k] = __ldcg(pl + (size_t) (((byp + dy) >> 1) + y) * cw + ((bxp + dx) >> 1) + x);
                }
            }
#pragma unroll
            for (int k = 0; k < 12; ++k)
                S.ipred[slot][r][lane + 32 * k] = (int16_t) (sp ? (d.sp_amount() ? evx_lerp_quarter(a[k], b[k]) : evx_lerp_half(a[k], b[k])) : a[k]);
        }
        // stale samples of the row below, macroblock column n-1 (read by the intra search of
        // columns n and n+1; the row below overwrites them only after we finished column n+1)
        if (n >= 1 && by + 1 < g.mbh)
        {
            int row = lane >> 1, half = lane & 1, x0 = (n - 1) * EVX_MB + 8 * half;
            uint4 v = __ldcg(reinterpret_cast<const uint4 *>(cur.y + (size_t) (py + 16 + row) * g.w + x0));
            *reinterpret_cast<uint4 *>(&S.wy[(64 + row) * EVX_RING_PWY + ((x0 >> 1) & 63)]) = v;
            if (lane < 16)
            {
                const int16_t *pl = lane < 8 ? cur.u : cur.v;
                int cx0 = (n - 1) * 8;
                uint4 c = __ldcg(reinterpret_cast<const uint4 *>(pl + (size_t) ((py >> 1) + 8 + (lane & 7)) * cw + cx0));
                *reinterpret_cast<uint4 *>(&(lane < 8 ? S.wu : S.wv)[(32 + (lane & 7)) * EVX_RING_PWC + ((cx0 >> 1) & 31)]) = c;
            }
        }
        __syncwarp();
        if (lane == 0) evx_mbar_arrive(&S.full[slot]);
        // a look at the next macroblock's stamps (the search usually is far ahead): one round trip less on its path
        stamps_seen = false;
        if (nref && p.fuse_k2 && n + 1 < g.mbw)
            stamps_seen = __all_sync(0xFFFFFFFFu, evx_ld_relaxed_u32(&p.inter[(size_t) min(lane, nref - 1) * nmb + mb + 1].stamp) == p.stamp) != 0;
    }
    for (int m = max(0, g.mbw - 2); m < g.mbw; ++m) publish(m);
#ifdef EVX_K3_LOADER_STATS
    if (p.prof && lane == 0) { p.prof[by * 10 + 6] = ls_own; p.prof[by * 10 + 7] = ls_stamp_cyc; p.prof[by * 10 + 8] = ls_own_cyc; }
#endif
}

// ---------------------------------------------------------------- column loader warp
//
// Column c of the three macroblock rows above enters the window once (c,by-1) is complete.
// Macroblock n may start when columns <= n+1 are in (fullb); it needs column n+2 only for
// candidates with x >= px+17, for the sub-pel taps there, and -- as the proof that (n+2,by-1)
// is complete -- before it overwrites samples the row above still reads (full2).  Two virtual
// columns past the right edge carry the last macroblocks' signals.

__device__ __forceinline__ void evx_k3_column_loader(EvxK3Smem &S, const EvxK3Params &p, int by, int lane)
{
    const EvxGeom g = p.g;
    const int cw = g.w >> 1, py = by * EVX_MB;
    const EvxPlanes cur = p.ring[(int) (p.frame_index % (uint32_t) p.R)];
    const int *progress = p.sync + 2;
    auto pull_column = [&](int col)
    {
        uint4 v[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            int c = lane + 32 * k;                 // 96 chunks: 48 rows x 2 halves
            int row = c >> 1, half = c & 1, y = py - 48 + row;
            v[k] = make_uint4(0, 0, 0, 0);
            if (y >= 0) v[k] = __ldcg(reinterpret_cast<const uint4 *>(cur.y + (size_t) y * g.w + col * EVX_MB + 8 * half));
        }
        uint4 cv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            int c = lane + 32 * k;                 // 48 chunks: 2 planes x 24 rows
            cv[k] = make_uint4(0, 0, 0, 0);
            if (c < 48)
            {
                int plane = c / 24, row = c % 24, y = (py >> 1) - 24 + row;
                if (y >= 0) cv[k] = __ldcg(reinterpret_cast<const uint4 *>((plane ? cur.v : cur.u) + (size_t) y * cw + col * 8));
            }
        }
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            int c = lane + 32 * k;
            int row = c >> 1, half = c & 1;
            *reinterpret_cast<uint4 *>(&S.wy[row * EVX_RING_PWY + (((col * EVX_MB + 8 * half) >> 1) & 63)]) = v[k];
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
        {
            int c = lane + 32 * k;
            if (c < 48)
            {
                int plane = c / 24, row = c % 24;
                *reinterpret_cast<uint4 *>(&(plane ? S.wv : S.wu)[row * EVX_RING_PWC + (((col * 8) >> 1) & 31)]) = cv[k];
            }
        }
    };
    for (int c = 0; c < g.mbw + 2; ++c)
    {
        // macroblock c-3 finished: the barrier phases of macroblocks c-1 / c-2 are free again and
        // nobody reads ring column c-8 any more
        if (c >= 3) evx_mbar_wait(&S.empty[(c - 3) & 1], (uint32_t) (((c - 3) >> 1) & 1));
        if (by > 0)
        {
            if (lane == 0) evx_wait_ge_far(progress + by - 1, min(c, g.mbw - 1) + 1, p.wait);
            __syncwarp();
            if (c < g.mbw) pull_column(c);
        }
        __syncwarp();
        if (lane == 0)
        {
            if (c >= 1 && c - 1 < g.mbw) evx_mbar_arrive(&S.fullb[(c - 1) & 1]);
            if (c >= 2) evx_mbar_arrive(&S.full2[(c - 2) & 1]);
        }
    }
}

// ---------------------------------------------------------------- compute warps

__device__ __forceinline__ bool evx_intra_legal(int x, int y, int px, int py, const EvxGeom &g)
{
    return !(y > py - EVX_MB && x > px - EVX_MB) && !(x < 0 || x > g.w - EVX_MB || y < 0 || y > g.h - EVX_MB);   // motion.cpp:238-248
}

// UNROLLED: the five search rounds and the 384-element passes fully unrolled -- each round's step, cell table and buffer
// index are then compile-time constants: 1.90 -> 1.73 -> 1.63 ms per 1080p frame ALONE on the device.  Next to other
// kernels (frames of a stream pipelined, many streams) the SMs' instruction supply is what limits the rows, and the
// rolled form is the faster one: 1 443 -> 1 654 frames/s in the six-slot pipeline for 4 % more latency alone.  So the
// kernel that has the device to itself (evx_wavefront<1>) is built unrolled, the one that shares it (<2>) rolled.
template <bool UNROLLED, int CW>
__device__ __forceinline__ void evx_k3_compute(EvxK3Smem &S, const EvxK3Params &p, int by, int tid)
{
    constexpr int CT = CW * 32, CPW = 8 / CW;      // compute threads; search cells (and sub-pel directions) per compute warp
    const int warp = tid >> 5, lane = tid & 31;
    const EvxGeom g = p.g;
    const int cw = g.w >> 1, py = by * EVX_MB;
    const int thr = (p.quality >> 2) + 1;                                  // motion.cpp:369, 436
    const int dest = (int) (p.frame_index % (uint32_t) p.R);               // common.cpp:192-195
    const EvxPlanes cur = p.ring[dest];
    const int nref = p.frame_type == 1 ? p.R - 1 : 0;
    EvxMbShared &sh = S.sh;
    int16_t *wy16 = reinterpret_cast<int16_t *>(S.wy), *wu16 = reinterpret_cast<int16_t *>(S.wu), *wv16 = reinterpret_cast<int16_t *>(S.wv);

    EvxRingWin win;
    win.y = S.wy; win.u = S.wu; win.v = S.wv; win.oy = py - 48; win.coy = (py >> 1) - 24;

    // this thread's row of the DCT basis (fixed output index i = tid & 7 in every pass)
    int lut_f[8], lut_i[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lut_f[k] = evx_dct_lut(tid & 7, k); lut_i[k] = evx_dct_lut(k, tid & 7); }

    // the search cells this warp owns: (dx,dy) in units of the round's step
    int cdx0[CPW], cdy0[CPW], cdx[CPW], cdy[CPW], ccell0[CPW], ccell[CPW];
#pragma unroll
    for (int q = 0; q < CPW; ++q)
    {
        const int k = warp + q * CW;
        ccell0[q] = k; cdx0[q] = k % 3 - 1; cdy0[q] = k / 3 - 2;          // round 0: rows -32,-16,0
        ccell[q] = k < 4 ? k : k + 1; cdx[q] = ccell[q] % 3 - 1; cdy[q] = ccell[q] / 3 - 1;
    }

    // selecting side: lane c owns cell c of a round (visiting order j-major, motion.cpp:232-233)
    const int lc = lane < 9 ? lane : 0, ldx = lc % 3 - 1, ldy = lc / 3;

    uint32_t n_full = 0, n_sub = 0;
    if (tid == 0) { S.last_motion = -1; S.last_coded = -1; }      // running, for K8; thread 0 only
    int row_records = 0;
    long long prof[10] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 }, tprev = clock64();
#ifdef EVX_K3_TRACE
#define EVX_K3_STAMP(k) do { if (p.prof && tid == 0) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); p.prof[(size_t) g.mbh * 10 + (size_t) mb * 4 + (k)] = (long long) gt_; } } while (0)
#else
#define EVX_K3_STAMP(k) do { } while (0)
#endif
#define EVX_K3_PROF(k) do { if (p.prof) { long long tn = clock64(); prof[k] += tn - tprev; tprev = tn; } } while (0)

    for (int n = 0; n < g.mbw; ++n)
    {
        const int slot = n & 1, px = n * EVX_MB, mb = by * g.mbw + n;
        evx_mbar_wait(&S.full[slot], (uint32_t) ((n >> 1) & 1));
#ifdef EVX_K3_STATS
        EVX_K3_PROF(9);              // waiting for the block loader (source, K2 results, predictions)
#endif
        evx_mbar_wait(&S.fullb[slot], (uint32_t) ((n >> 1) & 1));
        EVX_K3_PROF(0);
        EVX_K3_STAMP(0);
        bool have2 = false;      // far-right column (and write permission) not yet confirmed
#define EVX_K3_NEED2() do { if (!have2) { evx_mbar_wait(&S.full2[slot], (uint32_t) ((n >> 1) & 1)); have2 = true; EVX_K3_STAMP(1); } } while (0)
        const int16_t *srcb = S.src[slot];

        EvxLaneSrc src;
        {
            EvxLaneBlock sb;
            int rr = lane >> 3, cc = lane & 7;
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                int y = rr + 4 * k, x = 2 * cc;
                sb.w[k] = *reinterpret_cast<const uint32_t *>(&srcb[((y >> 3) * 2 + (x >> 3)) * 64 + (y & 7) * 8 + (x & 7)]);
            }
            int e = 256 + (lane >> 2) * 8 + 2 * (lane & 3);
            sb.w[4] = *reinterpret_cast<const uint32_t *>(&srcb[e]);
            sb.w[5] = *reinterpret_cast<const uint32_t *>(&srcb[e + 64]);
            evx_make_src(sb, src);
        }

        // ---- intra search, motion.cpp:354-419
        EvxSel s;
        s.bx = px; s.by = py; s.mad = EVX_BIG; s.ssd = EVX_BIG; s.sp_index = 0; s.sp_amount = 0; s.sp_enabled = 0;
        {   // compute_block_sad(src) with the int16 abs overload (analysis.h:57-68)
            int a = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += evx_abs16(evx_lo16(src.pos[k])) + evx_abs16(evx_hi16(src.pos[k]));
            s.sad = __reduce_add_sync(0xFFFFFFFFu, a);
        }
        int buf = 0;
#ifdef EVX_K3_STATS
        uint32_t n_hold = 0;
#endif
#pragma unroll (UNROLLED ? 5 : 1)
        for (int round = 0; round < 5; ++round)
        {
            const int step = EVX_SEARCH_RADIUS >> (round == 0 ? 0 : round);
            const int top = round == 0 ? -2 : -1;          // first round scans rows -32,-16,0 (motion.cpp:384-386)
            if (s.bx + step >= px + 17) EVX_K3_NEED2();      // a cell of this round reaches into column n+2
            // Round 0: grid cells 0..7 are evaluated (cells 7,8 = (0,0),(16,0) are never legal).
            // Later rounds: the centre cell 4 is the running best itself; its sad/mad/ssd are the
            // state's own (it was evaluated when it was accepted), so only the 8 outer cells run.
            // Every cell is evaluated unconditionally (any position a round can name lies inside the
            // staged window rows, and the ring masks the columns); illegal ones are flagged, not skipped,
            // which keeps the cells of a warp in one basic block.
            {
                // this warp's cells: loads and packed arithmetic first, then every reduction back to back;
                // only the raw (sad, mad) pair is published -- keys are built by the selecting lanes
                int la[CPW], lm[CPW];
#pragma unroll
                for (int q = 0; q < CPW; ++q)
                {
                    const int x = s.bx + (round == 0 ? cdx0[q] : cdx[q]) * step;
                    const int y = s.by + (round == 0 ? cdy0[q] : cdy[q]) * step;
                    EvxLaneBlock ref;
#ifdef EVX_K3_NO_ALIGNED_LOADS
                    evx_load_block_ring_bf(win, x, y, lane, ref);
#else
                    // (the loop is unrolled: `round` is a constant here)
                    if (round <= 2) evx_load_block_ring_al<false>(win, x, y, lane, ref);
                    else if (round == 3) evx_load_block_ring_al<true>(win, x, y, lane, ref);
                    else evx_load_block_ring_bf(win, x, y, lane, ref);
#endif
                    la[q] = evx_block_sad_lane(ref, src);
                    lm[q] = evx_block_mad_lane(ref, src);
                }
#pragma unroll
                for (int q = 0; q < CPW; ++q)
                {
                    const int sad = __reduce_add_sync(0xFFFFFFFFu, la[q]), mad = __reduce_max_sync(0xFFFFFFFFu, lm[q]);
                    if (lane == 0) S.cand2[buf][round == 0 ? ccell0[q] : ccell[q]] = make_int2(sad, mad);
                }
            }
            // selecting side, lane c = cell c in visiting order: everything that does not depend on the
            // published costs (position, legality, distance) is computed before the barrier
            const int cxl = s.bx + ldx * step, cyl = s.by + (top + ldy) * step;
            const int ssdl = (cxl - px) * (cxl - px) + (cyl - py) * (cyl - py);
            const bool from_state = round != 0 && lane == 4;            // the centre is the running best itself
            const bool legall = lane < 9 && evx_intra_legal(cxl, cyl, px, py, g) && !(round == 0 && lane == 8);
            evx_compute_sync<CT>();
            {
                int2 v = S.cand2[buf][lc];
                if (from_state) v = make_int2(s.sad, s.mad);
                const int wcell = evx_select_fullpel(s, evx_make_keys(v.x, v.y, ssdl, thr, legall), lane, thr, n_full);
#ifdef EVX_K3_STATS
                n_hold += wcell < 0 ? 1u << (6 * round) : 0u;      // five 6-bit counters... per macroblock, folded below
#endif
                if (wcell >= 0)
                {
                    s.bx = __shfl_sync(0xFFFFFFFFu, cxl, wcell); s.by = __shfl_sync(0xFFFFFFFFu, cyl, wcell);
                    s.sad = __shfl_sync(0xFFFFFFFFu, v.x, wcell); s.mad = __shfl_sync(0xFFFFFFFFu, v.y, wcell);
                    s.ssd = __shfl_sync(0xFFFFFFFFu, ssdl, wcell);
                }
            }
            buf ^= 1;
        }
#ifdef EVX_K3_STATS
        for (int r = 1; r < 5; ++r) prof[5 + r] += (n_hold >> (6 * r)) & 63u;      // round 0 always moves; slot 5 counts far-column waits
#endif
        EVX_K3_PROF(1);
        // ---- intra sub-pel, motion.cpp:277-317: eight directions, both blends each
        {
            if (s.bx + 1 >= px + 17) EVX_K3_NEED2();
#pragma unroll
            for (int q = 0; q < CPW; ++q)
            {
                const int k = warp + q * CW;                       // direction slot 0..7 (centre skipped)
                const int x = s.bx + cdx[q], y = s.by + cdy[q];
                EvxLaneBlock best, nb;
                int shh, mh, sq, mq;
                evx_load_block_ring_bf(win, s.bx, s.by, lane, best);
                evx_load_block_ring_bf(win, x, y, lane, nb);
                evx_subpel_cost(best, nb, src, -1, shh, mh, sq, mq);
                const int ok = evx_intra_legal(x, y, px, py, g) ? 1 : 0;
                if (lane == 0) { S.cand[buf][2 * k] = make_int4(shh, mh, 0, ok); S.cand[buf][2 * k + 1] = make_int4(sq, mq, 0, ok); }
            }
            evx_compute_sync<CT>();
            // sub-pel acceptance in closed form (evx_select_subpel), lane t = test t
            {
                const int4 v = lane < 16 ? S.cand[buf][lane] : make_int4(0, 0, 0, 0);
                const int wt = evx_select_subpel(s, v.x, v.y, v.w != 0, lane, thr, n_sub);
                if (wt >= 0)
                {
                    const int dd = (wt >> 1) < 4 ? (wt >> 1) : (wt >> 1) + 1;
                    s.sp_enabled = 1; s.sp_amount = wt & 1; s.sp_index = evx_frac_index(dd % 3 - 1, dd / 3 - 1);
                    s.sad = __shfl_sync(0xFFFFFFFFu, v.x, wt); s.mad = __shfl_sync(0xFFFFFFFFu, v.y, wt);
                }
            }
        }

        EVX_K3_PROF(2);
        // ---- classify, encode.cpp:17-67
        EvxDesc d = evx_desc_from_sel(s, 1, 0, px, py, thr);
        int best_sad = s.sad, best_ref = -1;
#pragma unroll 1
        for (int r = 0; r < nref; ++r)
        {
            const int4 raw = S.idesc[slot][r];
            const int isad = S.isad[slot][r];
            const bool cc = (raw.x & EVX_T_COPY) != 0, bc = (d.type() & EVX_T_COPY) != 0;
            const bool take = (cc != bc) ? cc : (isad < best_sad);
            if (take) { d.w0 = raw.x; d.w1 = raw.y; d.w2 = raw.z; d.w3 = raw.w; best_sad = isad; best_ref = r; }
        }
        const int type = d.type();
        const bool has_pred = type != EVX_T_INTRA;

        // ---- prediction (encode.cpp:83-141): intra from the window, inter from the loader's copy
        if (has_pred)
        {
            if (type & EVX_T_INTRA)
            {
                const int bxp = px + d.mx(), byp = py + d.my();
                const bool sp = d.sp_pred() != 0;
                int dx = 0, dy = 0;
                if (sp) evx_frac_direction(d.sp_index(), dx, dy);
                EVX_K3_FOR384(e)
                {
                    int comp, x, y, a, b = 0;
                    evx_mb_pos(e, comp, x, y);
                    if (comp == 0)
                    {
                        a = evx_ring_y(S, bxp + x, byp + y - win.oy);
                        if (sp) b = evx_ring_y(S, bxp + dx + x, byp + dy + y - win.oy);
                    }
                    else
                    {
                        const uint32_t *pl = comp == 1 ? S.wu : S.wv;
                        a = evx_ring_c(pl, (bxp >> 1) + x, (byp >> 1) + y - win.coy);
                        if (sp) b = evx_ring_c(pl, ((bxp + dx) >> 1) + x, ((byp + dy) >> 1) + y - win.coy);
                    }
                    sh.pred[e] = (int16_t) (sp ? (d.sp_amount() ? evx_lerp_quarter(a, b) : evx_lerp_half(a, b)) : a);
                }
            }
            else
            {
                EVX_K3_FOR384(e) sh.pred[e] = S.ipred[slot][best_ref][e];
            }
        }
        evx_compute_sync<CT>();
        EVX_K3_PROF(3);

        // Writing this macroblock needs no further wait.  The row above reads the stale samples under it only
        // through its own shared-memory copy, which its block loader makes while staging macroblock n+1; that
        // macroblock has completed (fullb, waited for above, needs progress[by-1] >= n+2), so the copy exists.
        // Column n+2 of the row above is awaited only by a search that actually reaches into it (EVX_K3_NEED2):
        // most macroblocks follow the row above at a lag of 2, not 3.
        // reconstruction target: global ring slot + our own window rows (py..py+15 -> 48..63)
        auto store_recon = [&](int e, int v)
        {
            int comp, x, y;
            evx_mb_pos(e, comp, x, y);
            if (comp == 0)
            {
                cur.y[(size_t) (py + y) * g.w + px + x] = (int16_t) v;
                wy16[(48 + y) * (EVX_RING_PWY * 2) + ((px + x) & 127)] = (int16_t) v;
            }
            else
            {
                (comp == 1 ? cur.u : cur.v)[(size_t) ((py >> 1) + y) * cw + (px >> 1) + x] = (int16_t) v;
                (comp == 1 ? wu16 : wv16)[(24 + y) * (EVX_RING_PWC * 2) + (((px >> 1) + x) & 63)] = (int16_t) v;
            }
        };

        if (type & EVX_T_COPY)
        {   // copy blocks: the prediction is the reconstruction; no coefficients (encode.cpp:155-157)
            EVX_K3_FOR384(e) store_recon(e, sh.pred[e]);
            if (tid == 0) { p.table[mb] = d; p.prev_motion[mb] = S.last_motion; p.prev_coded[mb] = S.last_coded; }
        }
        else
        {
            // The four 1-D passes keep one output per thread (index i = tid & 7, whose basis row sits in registers)
            // but fetch their eight inputs with ONE 16-byte shared-memory load: the passes that walk columns read
            // buffers their producers wrote transposed.  (Sixteen 2-byte loads per output made every pass
            // load-bound: ~800 cycles each; a non-copy macroblock is the slow case of the wavefront, and the spread
            // between fast and slow macroblocks costs the frame as much as their mean.)
            auto unpack8 = [](const uint4 &v, int x[8])
            {
                x[0] = evx_lo16(v.x); x[1] = evx_hi16(v.x); x[2] = evx_lo16(v.y); x[3] = evx_hi16(v.y);
                x[4] = evx_lo16(v.z); x[5] = evx_hi16(v.z); x[6] = evx_lo16(v.w); x[7] = evx_hi16(v.w);
            };
            // residual (int16, transform.cpp:29-32) and row pass (transform.cpp:264-301); result stored transposed
            EVX_K3_FOR384(e)
            {
                const int base = e & ~7;
                uint4 sv = *reinterpret_cast<const uint4 *>(srcb + base);
                if (has_pred)
                {
                    const uint4 pv = *reinterpret_cast<const uint4 *>(sh.pred + base);
                    sv.x = __vsub2(sv.x, pv.x); sv.y = __vsub2(sv.y, pv.y); sv.z = __vsub2(sv.z, pv.z); sv.w = __vsub2(sv.w, pv.w);   // wraps like the int16 store
                }
                int x[8];
                unpack8(sv, x);
                int t = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += x[k] * lut_f[k];
                t = (e & 7) == 0 ? evx_tdiv_pow2(t * 45, 7) : evx_tdiv_pow2(t, 1);
                sh.bufb[(e & ~63) + (e & 7) * 8 + ((e >> 3) & 7)] = (int16_t) evx_rdiv_pow2(t, 7);      // [block][i][row]
            }
            evx_compute_sync<CT>();
            // column pass: thread e -> (block b, column a, output row i = e & 7); column a is contiguous in bufb
            uint32_t vsum = 0, vsq = 0; int vcnt = 0;
            EVX_K3_FOR384(e)
            {
                const int b = e >> 6, a = (e >> 3) & 7, i = e & 7;
                int x[8];
                unpack8(*reinterpret_cast<const uint4 *>(sh.bufb + b * 64 + a * 8), x);
                int t = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) t += x[k] * lut_f[k];
                t = i == 0 ? evx_tdiv_pow2(t * 45, 7) : evx_tdiv_pow2(t, 1);
                const int16_t c16 = (int16_t) evx_rdiv_pow2(t, 7);
                sh.bufa[b * 64 + i * 8 + a] = c16;
#ifndef EVX_K3_VAR_PASS
                // compute_block_variance2 over the 16x16 luma coefficients except (0,0) (analysis.h:176-198): the sums are
                // taken here, from the value each thread just produced (uint32 sums: any order gives the same bits)
                if (b < 4 && (b | i | a) != 0 && c16 != 0) { vsum += (uint32_t) (int) c16; vsq += (uint32_t) ((int) c16 * (int) c16); vcnt++; }
#endif
            }
#ifndef EVX_K3_VAR_PASS
            {
                const uint32_t sum = __reduce_add_sync(0xFFFFFFFFu, vsum), sq = __reduce_add_sync(0xFFFFFFFFu, vsq);
                const int cnt = __reduce_add_sync(0xFFFFFFFFu, vcnt);
                if (lane == 0) { sh.red[warp] = (int) sum; sh.red[32 + warp] = (int) sq; sh.red[64 + warp] = cnt; }
            }
            evx_compute_sync<CT>();
#else
            evx_compute_sync<CT>();
            // compute_block_variance2 over the 16x16 luma coefficients except (0,0) (analysis.h:176-198)
            {
                uint32_t sum = 0, sq = 0; int cnt = 0;
                for (int e = tid; e < 256; e += CT) { int t = e ? sh.bufa[e] : 0; if (t) { sum += (uint32_t) t; sq += (uint32_t) (t * t); cnt++; } }
                sum = __reduce_add_sync(0xFFFFFFFFu, sum); sq = __reduce_add_sync(0xFFFFFFFFu, sq); cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
                if (lane == 0) { sh.red[warp] = (int) sum; sh.red[32 + warp] = (int) sq; sh.red[64 + warp] = cnt; }
                evx_compute_sync<CT>();
            }
#endif
            int qp, var = 0;
            {
                uint32_t Ssum = 0, Q = 0; int C = 0;
#pragma unroll
                for (int w8 = 0; w8 < CW; ++w8) { Ssum += (uint32_t) sh.red[w8]; Q += (uint32_t) sh.red[32 + w8]; C += sh.red[64 + w8]; }
                if (C > 0) var = (int) (Q - (uint32_t) evx_rdiv((int) (Ssum * Ssum), C));
                // query_block_quantization_parameter, quantize.cpp:60-77
                const int q = p.quality & 0xFF;
                const int idx = evx_clip(evx_ilog2((uint32_t) var) >> 1, 1, 31);
                qp = q;
                if (idx > q) qp = evx_clip(q + ((idx - q) >> 1), 1, 31);
                else if (idx < q) qp = evx_clip(q - ((q - idx) >> 1), 1, 31);
            }
            const bool intra_q = (type & EVX_T_INTRA) && !(type & EVX_T_MOTION);
            int16_t *rec = p.records + (size_t) mb * 384;
            // quantise (quantize.cpp:79-180), hand the record out, dequantise (quantize.cpp:182-243)
            EVX_K3_FOR384(e)
            {
                const int mode = intra_q ? ((e >> 6) < 4 ? 0 : 1) : 2;
                const int qv = evx_quant_fast(sh.bufa[e], e & 63, mode, qp, p.linear, sh.qmi, sh.qmt, sh.recip);
                rec[evx_record_index(e)] = (int16_t) qv;
                sh.bufb[(e & ~63) + (e & 7) * 8 + ((e >> 3) & 7)] = (int16_t) evx_dequant(qv, e & 63, mode, qp, p.linear, sh.qmi, sh.qmt);   // [block][column][row]
            }
            if (tid == 0) { d.set_q(qp, var); p.table[mb] = d; p.prev_motion[mb] = S.last_motion; p.prev_coded[mb] = S.last_coded; S.last_coded = mb; }
            evx_compute_sync<CT>();
            // inverse transform (transform.cpp:330-366, 418-433): columns, then rows + prediction
            EVX_K3_FOR384(e)
            {
                const int b = e >> 6, j = (e >> 3) & 7;      // column j (contiguous in bufb), output row i = e & 7
                int x[8];
                unpack8(*reinterpret_cast<const uint4 *>(sh.bufb + b * 64 + j * 8), x);
                int t = evx_tdiv_pow2((x[0] * lut_i[0]) * 45, 7);
#pragma unroll
                for (int k = 1; k < 8; ++k) t += evx_tdiv_pow2(x[k] * lut_i[k], 1);
                sh.bufa[b * 64 + (e & 7) * 8 + j] = (int16_t) evx_rdiv_pow2(t, 7);
            }
            evx_compute_sync<CT>();
            EVX_K3_FOR384(e)
            {
                int x[8];                                    // row (e>>3), output column i = e & 7
                unpack8(*reinterpret_cast<const uint4 *>(sh.bufa + (e & ~7)), x);
                int t = evx_tdiv_pow2((x[0] * lut_i[0]) * 45, 7);
#pragma unroll
                for (int k = 1; k < 8; ++k) t += evx_tdiv_pow2(x[k] * lut_i[k], 1);
                int v = evx_rdiv_pow2(t, 7);
                if (has_pred) v += sh.pred[e];
                store_recon(e, v);
            }
            row_records++;
        }
#ifdef EVX_K3_STATS
        prof[5] += have2 ? 1 : 0;       // macroblocks that waited for column n+2 of the row above
#endif
        if (tid == 0 && (type & EVX_T_MOTION)) S.last_motion = mb;
        evx_compute_sync<CT>();
        EVX_K3_PROF(4);
        EVX_K3_STAMP(2);
        if (tid == 0)
        {
            evx_mbar_arrive(&S.empty[slot]);
#ifdef EVX_K3_TIMELINE
            if (p.prof && n == 0) p.prof[(size_t) g.mbh * 10 + (size_t) by * 4 + 3] = (long long) evx_globaltimer();     // first macroblock of the row complete
#endif
        }
    }
    if (tid == 0)
    {
        p.row_records[by] = row_records;
        p.row_last[by] = S.last_motion; p.row_last[g.mbh + by] = S.last_coded;
#ifdef EVX_K3_LOADER_STATS
        if (p.prof) for (int k = 0; k < 10; ++k) if (k < 6 || k > 8) p.prof[by * 10 + k] = prof[k];      // (6..8: the block loader's)
#else
        if (p.prof) for (int k = 0; k < 10; ++k) p.prof[by * 10 + k] = prof[k];
#endif
        atomicAdd(&p.counters[2], (unsigned long long) n_full);
        atomicAdd(&p.counters[3], (unsigned long long) n_sub);
    }
}

// Two register budgets of the same kernel: <2> keeps two CTAs per SM (96 registers) -- what several frames or streams
// sharing the device need; <1> lets ptxas have the 168 registers it asks for, 2.5 % faster per frame alone
// (1.735 -> 1.691 ms at 1080p), used for a frame that has the device to itself (the stand-alone launch sequence).
template <int MINCTAS>
__global__ void __launch_bounds__(EvxK3Cfg<MINCTAS>::NT, MINCTAS) evx_wavefront(const __grid_constant__ EvxK3Params p)
{
    extern __shared__ __align__(128) uint8_t evx_k3_smem[];
    EvxFrameCtl &C = *reinterpret_cast<EvxFrameCtl *>(evx_k3_smem);
    EvxK3Smem &S = *reinterpret_cast<EvxK3Smem *>(evx_k3_smem + EVX_FRAME_CTL_BYTES);
    constexpr int CW = EvxK3Cfg<MINCTAS>::CW;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.g.mbh;

    if (tid == 0 && p.started) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p.started), "r"(p.dbk_base) : "memory");
    // Persistent over tickets = wavefront rows, claimed in order: row r can only be active during macroblock steps
    // [3r, 3r + W), so at most ceil(W/3) rows are in flight at any time; the host sizes the grid for that many CTAs plus
    // the ones that serve the search queue while they wait for their row.  A CTA only ever waits on rows claimed before
    // its own (by CTAs that are running), on search items that are claimed, and on the previous frame.
    auto inval = [](uint64_t *bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(evx_smem_addr(bar)) : "memory"); };
    for (;;)
    {
        if (tid == 0) C.row = atomicAdd(&p.sync[0], 1);
        __syncthreads();
        const int by = C.row;
        if (by >= H) return;
#ifdef EVX_K3_TIMELINE
        if (p.prof && tid == 0) p.prof[(size_t) H * 10 + (size_t) by * 4 + 0] = (long long) evx_globaltimer();      // row claimed
#endif
        if (tid == 0)
        {
            evx_mbar_init(&S.full[0], 1); evx_mbar_init(&S.full[1], 1); evx_mbar_init(&S.fullb[0], 1); evx_mbar_init(&S.fullb[1], 1);
            evx_mbar_init(&S.full2[0], 1); evx_mbar_init(&S.full2[1], 1); evx_mbar_init(&S.empty[0], 1); evx_mbar_init(&S.empty[1], 1);
            evx_mbar_init(&S.k2bar, 1);
        }
        evx_init_tables(S.sh, tid, EvxK3Cfg<MINCTAS>::NT);
        __syncthreads();
#ifdef EVX_K3_TIMELINE
        if (p.prof && tid == 0) p.prof[(size_t) H * 10 + (size_t) by * 4 + 1] = (long long) evx_globaltimer();      // left the queues, row starts
#endif
#ifdef EVX_K3_ALONE_ROLLED
        if (warp < CW) evx_k3_compute<false, CW>(S, p, by, tid);
#else
        if (warp < CW) evx_k3_compute<MINCTAS == 1, CW>(S, p, by, tid);
#endif
        else if (warp == CW) evx_k3_block_loader(S, p, by, lane);
        else if (warp == CW + 1) evx_k3_column_loader(S, p, by, lane);
        __syncthreads();      // every warp has left the row: its barriers and shared memory may be reused
#ifdef EVX_K3_TIMELINE
        if (p.prof && tid == 0) p.prof[(size_t) H * 10 + (size_t) by * 4 + 2] = (long long) evx_globaltimer();      // row complete
#endif
        if (tid == 0)
        {
            inval(&S.full[0]); inval(&S.full[1]); inval(&S.fullb[0]); inval(&S.fullb[1]);
            inval(&S.full2[0]); inval(&S.full2[1]); inval(&S.empty[0]); inval(&S.empty[1]); inval(&S.k2bar);
        }
        if (p.fuse_dbk && by == H - 1)
        {   // the frame's last row is complete, so every deblocking job is ready: whatever nobody has taken yet is done here
            int r;
            while ((r = evx_serve_deblock(p, lane, (unsigned int) warp * 7u)) >= 0) if (r == 0) __nanosleep(200);
        }
    }
}

// ------------------------------------------------------------------ the followers (header comment)

#ifndef EVX_SF_WARPS
#define EVX_SF_WARPS 8            // warps of a search-follower CTA: a row's 120 items in ~300 us, twice the pace its wavefront row needs
#endif
struct EvxSearchFollowSmem
{
    uint8_t win[EVX_SF_WARPS][EVX_K2W_BYTES];
    uint64_t bar[EVX_SF_WARPS];
    int row;
};

__global__ void __launch_bounds__(EVX_SF_WARPS * 32, 4) evx_search_follow(const __grid_constant__ EvxK3Params p)
{
    extern __shared__ __align__(128) uint8_t evx_sf_smem[];
    EvxSearchFollowSmem &F = *reinterpret_cast<EvxSearchFollowSmem *>(evx_sf_smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.g.mbh, nref = p.R - 1, per_row = p.g.mbw * nref;
    int *ticket = p.sync + 2 + 3 * H;
    if (tid == 0) for (int w = 0; w < EVX_SF_WARPS; ++w) evx_mbar_init(&F.bar[w], 1);
    uint32_t phase = 0;
    for (;;)
    {
        __syncthreads();
        if (tid == 0) F.row = atomicAdd(ticket, 1);
        __syncthreads();
        const int y = F.row;
        if (y >= H) return;
        int *k2c = evx_k2c(p) + y;
        for (;;)
        {
            int it = 0;
            if (lane == 0) it = atomicAdd(k2c, 1);
            it = __shfl_sync(0xFFFFFFFFu, it, 0);
            if (it >= per_row) break;
            evx_gate_prev(p, it / nref, y, lane);
            evx_run_search_item(p, y, it, lane, F.win[warp], &F.bar[warp], phase);
        }
    }
}

__global__ void __launch_bounds__(32, 16) evx_deblock_follow(const __grid_constant__ EvxK3Params p)
{
    const int lane = threadIdx.x;
    const int H = p.g.mbh, W = p.g.mbw, nch = (W + EVX_DBK_CHUNK - 1) / EVX_DBK_CHUNK;
    int *ticket = p.sync + 3 + 3 * H;
    const int *progress = evx_progress(p);
    for (;;)
    {
        int Y = 0;
        if (lane == 0) Y = atomicAdd(ticket, 1);
        Y = __shfl_sync(0xFFFFFFFFu, Y, 0);
        if (Y >= H) return;
        int *jc = evx_jc(p) + Y;
        const int *prow = progress + min(Y + 3, H - 1);
        for (;;)
        {
            // the row's next job, once the wavefront has passed it: macroblock min(4c+5, W-1) of row min(Y+3, H-1) complete
            int c = 0;
            if (lane == 0) c = evx_ld_relaxed(jc);
            c = __shfl_sync(0xFFFFFFFFu, c, 0);
            if (c >= nch) break;
            const int cnt = min(c * EVX_DBK_CHUNK + EVX_DBK_CHUNK, W), need = min(cnt + 1, W - 1) + 1;
            if (lane == 0)
            {
                int have;
                EVX_BOUNDED_WAIT(p.wait, (have = evx_ld_relaxed(prow)) >= need, min(EVX_DBK_POLL_NS, 500 * (need - have)), 9u, (unsigned int) Y, (unsigned int) need, (unsigned int) have);
            }
            __syncwarp();
            int ok = 0;
            if (lane == 0) ok = atomicCAS(jc, c, c + 1) == c;
            if (!__shfl_sync(0xFFFFFFFFu, ok, 0)) continue;           // (the wavefront kernel's last row took it)
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            evx_deblock_job(p, Y, c * EVX_DBK_CHUNK, cnt, lane);
        }
    }
}
