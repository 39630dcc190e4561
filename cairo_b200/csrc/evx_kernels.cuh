// evx_kernels.cuh -- the sm_100a kernels of the EVX-1 pixel pipeline.
//
//   K1 evx_rgb_to_yuv420      convert.cpp:95-160      HBM bound, element-wise
//   K2 evx_inter_search       motion.cpp:421-494      integer-ALU bound; TMA-staged windows
//   K3 evx_wavefront          motion.cpp:354-419 + encode.cpp:17-203 + decode.cpp:15-144
//   K4 evx_deblock            deblock.cpp:201-275     HBM bound; independent 8x8 tiles
//   K5 evx_decode_recon       decode.cpp:146-170
//   K6 evx_yuv420_to_rgb      convert.cpp:162-223     HBM bound, element-wise
#pragma once

#include <cuda.h>

#include "evx_device.cuh"

// the arithmetic of one 8x2 strip (convert.cpp:11-14, 30-73)
__device__ __forceinline__ void evx_rgb8x2_to_yuv(const uint8_t (&px)[2][24], short (&yv)[2][8], short (&uv)[4], short (&vv)[4])
{
#pragma unroll
    for (int q = 0; q < 4; ++q)
    {
        short su = 0, sv = 0;
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int d = 0; d < 2; ++d)
        {
            int i = 2 * q + d;
            int R = px[r][3 * i], G = px[r][3 * i + 1], B = px[r][3 * i + 2];
            yv[r][i] = (short) (((77 * R + 150 * G + 29 * B + 128) >> 8) + 16);
            su = (short) (su + ((-43 * R - 85 * G + 128 * B + 128) / 256 + 128));
            sv = (short) (sv + ((128 * R - 107 * G - 21 * B + 128) / 256 + 128));
        }
        uv[q] = (short) ((su + 2) >> 2);
        vv[q] = (short) ((sv + 2) >> 2);
    }
}

// ------------------------------------------------------------------ K1: RGB -> YUV 4:2:0
// One thread per 8x2 pixel strip: 2 x 24 B in, 2 x 16 B luma + 2 x 8 B chroma out.
// y uses >>8, chroma uses C '/' (toward zero); the four chroma samples of a quad are
// summed in an int16 and averaged with (sum+2)>>2 (convert.cpp:11-14, 30-73).
// Samples outside the visible frame are never written (they stay 0).
__global__ void __launch_bounds__(256) evx_rgb_to_yuv420(const uint8_t *__restrict__ rgb, EvxPlanes dst, EvxGeom g)
{
    int strips_x = (g.vw + 7) >> 3;
    int sx = blockIdx.x * blockDim.x + threadIdx.x;
    int sy = blockIdx.y;
    if (sx >= strips_x) return;
    int x0 = sx * 8, y0 = sy * 2;
    int n = min(8, g.vw - x0);                       // visible pixels in this strip (even)
    uint8_t px[2][24];
    bool fast = (n == 8) && ((g.vw & 3) == 0) && ((((size_t) rgb) & 3) == 0);
#pragma unroll
    for (int r = 0; r < 2; ++r)
    {
        const uint8_t *row = rgb + ((size_t) (y0 + r) * g.vw + x0) * 3;
        if (fast)
        {
            const uint32_t *row4 = reinterpret_cast<const uint32_t *>(row);
#pragma unroll
            for (int k = 0; k < 6; ++k) { uint32_t v = __ldg(row4 + k); px[r][4 * k] = v; px[r][4 * k + 1] = v >> 8; px[r][4 * k + 2] = v >> 16; px[r][4 * k + 3] = v >> 24; }
        }
        else
        {
            for (int k = 0; k < 24; ++k) px[r][k] = k < n * 3 ? __ldg(row + k) : 0;
        }
    }
    short yv[2][8], uv[4], vv[4];
    evx_rgb8x2_to_yuv(px, yv, uv, vv);
    int cw = g.w >> 1;
    if (n == 8)
    {
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
            uint4 o;
            o.x = evx_pack16(yv[r][0], yv[r][1]); o.y = evx_pack16(yv[r][2], yv[r][3]);
            o.z = evx_pack16(yv[r][4], yv[r][5]); o.w = evx_pack16(yv[r][6], yv[r][7]);
            *reinterpret_cast<uint4 *>(dst.y + (size_t) (y0 + r) * g.w + x0) = o;
        }
        uint2 ou = { evx_pack16(uv[0], uv[1]), evx_pack16(uv[2], uv[3]) };
        uint2 ov = { evx_pack16(vv[0], vv[1]), evx_pack16(vv[2], vv[3]) };
        *reinterpret_cast<uint2 *>(dst.u + (size_t) sy * cw + (x0 >> 1)) = ou;
        *reinterpret_cast<uint2 *>(dst.v + (size_t) sy * cw + (x0 >> 1)) = ov;
    }
    else
    {
        for (int i = 0; i < n; ++i) { dst.y[(size_t) y0 * g.w + x0 + i] = yv[0][i]; dst.y[(size_t) (y0 + 1) * g.w + x0 + i] = yv[1][i]; }
        for (int q = 0; q < n / 2; ++q) { dst.u[(size_t) sy * cw + (x0 >> 1) + q] = uv[q]; dst.v[(size_t) sy * cw + (x0 >> 1) + q] = vv[q]; }
    }
}

// The same for frames whose width is a multiple of 16 and whose rows are 16-byte aligned: one thread per 16x2 strip, 2 x 3
// 128-bit loads in, 2 x 2 128-bit luma stores + 2 x 1 128-bit chroma stores out; threads flattened over the whole frame.
__global__ void __launch_bounds__(256) evx_rgb_to_yuv420_wide(const uint8_t *__restrict__ rgb, EvxPlanes dst, EvxGeom g)
{
    const int strips_x = g.vw >> 4;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= strips_x * (g.vh >> 1)) return;
    const int sy = id / strips_x, sx = id - sy * strips_x, x0 = sx * 16, y0 = sy * 2;
    uint4 in[2][3];
#pragma unroll
    for (int r = 0; r < 2; ++r)
    {
        const uint4 *row = reinterpret_cast<const uint4 *>(rgb + ((size_t) (y0 + r) * g.vw + x0) * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) in[r][k] = __ldg(row + k);
    }
    const int cw = g.w >> 1;
    uint4 ou, ov;
#pragma unroll
    for (int half = 0; half < 2; ++half)
    {
        uint8_t px[2][24];
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
            const uint32_t w[12] = { in[r][0].x, in[r][0].y, in[r][0].z, in[r][0].w, in[r][1].x, in[r][1].y, in[r][1].z, in[r][1].w, in[r][2].x, in[r][2].y, in[r][2].z, in[r][2].w };
#pragma unroll
            for (int k = 0; k < 6; ++k) { const uint32_t v = w[6 * half + k]; px[r][4 * k] = v; px[r][4 * k + 1] = v >> 8; px[r][4 * k + 2] = v >> 16; px[r][4 * k + 3] = v >> 24; }
        }
        short yv[2][8], uv[4], vv[4];
        evx_rgb8x2_to_yuv(px, yv, uv, vv);
#pragma unroll
        for (int r = 0; r < 2; ++r)
        {
            uint4 o;
            o.x = evx_pack16(yv[r][0], yv[r][1]); o.y = evx_pack16(yv[r][2], yv[r][3]);
            o.z = evx_pack16(yv[r][4], yv[r][5]); o.w = evx_pack16(yv[r][6], yv[r][7]);
            *reinterpret_cast<uint4 *>(dst.y + (size_t) (y0 + r) * g.w + x0 + 8 * half) = o;
        }
        if (half == 0) { ou.x = evx_pack16(uv[0], uv[1]); ou.y = evx_pack16(uv[2], uv[3]); ov.x = evx_pack16(vv[0], vv[1]); ov.y = evx_pack16(vv[2], vv[3]); }
        else { ou.z = evx_pack16(uv[0], uv[1]); ou.w = evx_pack16(uv[2], uv[3]); ov.z = evx_pack16(vv[0], vv[1]); ov.w = evx_pack16(vv[2], vv[3]); }
    }
    *reinterpret_cast<uint4 *>(dst.u + (size_t) sy * cw + (x0 >> 1)) = ou;
    *reinterpret_cast<uint4 *>(dst.v + (size_t) sy * cw + (x0 >> 1)) = ov;
}

// ------------------------------------------------------------------ K6: YUV 4:2:0 -> RGB
// convert.cpp:16-19, 162-223.  saturate() funnels through an int16 parameter (math.h:218-221).
__device__ __forceinline__ uint8_t evx_sat8(int v) { return (uint8_t) evx_clip((int) (short) v, 0, 255); }

__global__ void __launch_bounds__(256) evx_yuv420_to_rgb(EvxPlanes src, uint8_t *__restrict__ rgb, EvxGeom g)
{
    int strips_x = (g.vw + 7) >> 3;
    int sx = blockIdx.x * blockDim.x + threadIdx.x;
    int sy = blockIdx.y;
    if (sx >= strips_x) return;
    int x0 = sx * 8, y0 = sy * 2;
    int n = min(8, g.vw - x0);
    int cw = g.w >> 1;
    short uv[4], vv[4];
    {
        uint2 a = *reinterpret_cast<const uint2 *>(src.u + (size_t) sy * cw + (x0 >> 1));
        uint2 b = *reinterpret_cast<const uint2 *>(src.v + (size_t) sy * cw + (x0 >> 1));
        uv[0] = a.x; uv[1] = a.x >> 16; uv[2] = a.y; uv[3] = a.y >> 16;
        vv[0] = b.x; vv[1] = b.x >> 16; vv[2] = b.y; vv[3] = b.y >> 16;
    }
    bool fast = (n == 8) && ((g.vw & 3) == 0) && ((((size_t) rgb) & 3) == 0);
#pragma unroll
    for (int r = 0; r < 2; ++r)
    {
        uint4 yy = *reinterpret_cast<const uint4 *>(src.y + (size_t) (y0 + r) * g.w + x0);
        short yv[8] = { (short) yy.x, (short) (yy.x >> 16), (short) yy.y, (short) (yy.y >> 16), (short) yy.z, (short) (yy.z >> 16), (short) yy.w, (short) (yy.w >> 16) };
        uint8_t o[24];
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            int y = yv[i], u = uv[i >> 1], v = vv[i >> 1];
            o[3 * i]     = evx_sat8((256 * (y - 16) + 358 * (v - 128) + 128) >> 8);
            o[3 * i + 1] = evx_sat8((256 * (y - 16) - 88 * (u - 128) - 182 * (v - 128) + 128) >> 8);
            o[3 * i + 2] = evx_sat8((256 * (y - 16) + 452 * (u - 128) + 128) >> 8);
        }
        uint8_t *row = rgb + ((size_t) (y0 + r) * g.vw + x0) * 3;
        if (fast)
        {
            uint32_t *row4 = reinterpret_cast<uint32_t *>(row);
#pragma unroll
            for (int k = 0; k < 6; ++k) row4[k] = o[4 * k] | (o[4 * k + 1] << 8) | (o[4 * k + 2] << 16) | ((uint32_t) o[4 * k + 3] << 24);
        }
        else
        {
            for (int k = 0; k < n * 3; ++k) row[k] = o[k];
        }
    }
}

// The same for frames whose width is a multiple of 16 and whose rows are 16-byte aligned: one thread per 16x2 strip, 128-bit
// loads of both planes' rows, 2 x 3 128-bit stores.
__global__ void __launch_bounds__(256) evx_yuv420_to_rgb_wide(EvxPlanes src, uint8_t *__restrict__ rgb, EvxGeom g)
{
    const int strips_x = g.vw >> 4;
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= strips_x * (g.vh >> 1)) return;
    const int sy = id / strips_x, sx = id - sy * strips_x, x0 = sx * 16, y0 = sy * 2;
    const int cw = g.w >> 1;
    const uint4 ua = *reinterpret_cast<const uint4 *>(src.u + (size_t) sy * cw + (x0 >> 1));
    const uint4 va = *reinterpret_cast<const uint4 *>(src.v + (size_t) sy * cw + (x0 >> 1));
    const uint32_t uw[4] = { ua.x, ua.y, ua.z, ua.w }, vw4[4] = { va.x, va.y, va.z, va.w };
#pragma unroll
    for (int r = 0; r < 2; ++r)
    {
        const uint4 *yrow = reinterpret_cast<const uint4 *>(src.y + (size_t) (y0 + r) * g.w + x0);
        const uint4 ya = yrow[0], yb = yrow[1];
        const uint32_t yw[8] = { ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w };
        uint8_t o[48];
#pragma unroll
        for (int i = 0; i < 16; ++i)
        {
            const int y = (short) (yw[i >> 1] >> (16 * (i & 1))), u = (short) (uw[i >> 2] >> (16 * ((i >> 1) & 1))), v = (short) (vw4[i >> 2] >> (16 * ((i >> 1) & 1)));
            o[3 * i]     = evx_sat8((256 * (y - 16) + 358 * (v - 128) + 128) >> 8);
            o[3 * i + 1] = evx_sat8((256 * (y - 16) - 88 * (u - 128) - 182 * (v - 128) + 128) >> 8);
            o[3 * i + 2] = evx_sat8((256 * (y - 16) + 452 * (u - 128) + 128) >> 8);
        }
        uint4 *row = reinterpret_cast<uint4 *>(rgb + ((size_t) (y0 + r) * g.vw + x0) * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const int b = 16 * k + 4 * j; w[j] = o[b] | (o[b + 1] << 8) | (o[b + 2] << 16) | ((uint32_t) o[b + 3] << 24); }
            row[k] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

// ------------------------------------------------------------------ TMA / mbarrier (sm_100a PTX)

__device__ __forceinline__ uint32_t evx_smem_addr(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void evx_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(evx_smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void evx_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(evx_smem_addr(bar)), "r"(bytes) : "memory");
}

// try_wait with a suspend-time hint: the warp is parked by the hardware until the phase completes (or
// the hint runs out) instead of spinning through issue slots that other warps of the SM could use.
__device__ __forceinline__ void evx_mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(evx_smem_addr(bar)), "r"(phase), "r"(0x989680u) : "memory");
}

__device__ __forceinline__ void evx_tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(evx_smem_addr(dst)), "l"(map), "r"(x), "r"(y), "r"(evx_smem_addr(bar)) : "memory");
}

// ------------------------------------------------------------------ K2: inter search
//
// Each WARP runs the whole sequential search of one (macroblock, reference) pair (motion.cpp:421-494) with no
// block-level synchronisation: the 16x16+8x8+8x8 candidate cost is a warp collective (evx_block_cost), the
// acceptance rule has a closed form evaluated lane-parallel (evx_select_fullpel).

// 32 bytes per (macroblock, reference).  `stamp` is the token of the frame the record belongs to: when the search runs in
// the frame pipeline (evx_search_follow, evx_wavefront.cuh) it is stored last, with release semantics, and the wavefront's block loader
// polls it -- the slot's records are reused by later frames and never zeroed.
struct EvxInterResult { EvxDesc desc; int sad; uint32_t stamp; int pad[2]; };

struct EvxK2Maps { CUtensorMap m[3 * 7]; };   // [ref][Y,U,V] for up to 7 past references

__device__ __forceinline__ void evx_load_src_lane(const EvxPlanes &srcp, const EvxGeom &g, int px, int py, int lane, EvxLaneBlock &b)
{
    const uint32_t *y = reinterpret_cast<const uint32_t *>(srcp.y + (size_t) (py + (lane >> 3)) * g.w + px) + (lane & 7);
#pragma unroll
    // (L2 loads throughout: with frames overlapping on the device these planes are written by kernels that run at the
    // same time as the reader, which rules out the non-coherent path and a stale L1 line)
    for (int k = 0; k < 4; ++k) b.w[k] = __ldcg(y + (size_t) 4 * k * (g.w >> 1));
    int cw = g.w >> 1;
    size_t off = (size_t) ((py >> 1) + (lane >> 2)) * cw + (px >> 1);
    b.w[4] = __ldcg(reinterpret_cast<const uint32_t *>(srcp.u + off) + (lane & 3));
    b.w[5] = __ldcg(reinterpret_cast<const uint32_t *>(srcp.v + off) + (lane & 3));
}

// the sequential search of one macroblock against one window (motion.cpp:254-275, 319-352, 421-494).
// `centre` is the co-located block; `stage(k, cx, cy, win)` is called before the first round (k = 0, centre
// of the step-16 round) and after it (k = 1, its winner): a kernel whose window does not hold the whole
// +-32 range re-centres it there (every later position stays within [-16, +32) of that winner).
struct EvxNoStage { __device__ __forceinline__ void operator()(int, int, int, EvxWin &) const {} };

// CELL_UNROLL: how many of a round's eight cells are costed per loop iteration.  8 (the round as one basic block) is the
// fastest form for the stand-alone kernel (49 against 53 us per 1080p frame); 2 is a third of the code, which is what counts
// for the search follower next to the wavefront rows (evx_wavefront.cuh).
template <int CELL_UNROLL, class Stage>
__device__ __forceinline__ void evx_inter_search_warp(EvxWin &win, const EvxLaneSrc &src, const EvxLaneBlock &centre, const EvxGeom &g, int px, int py, int thr,
                                                      int lane, EvxSel &s, uint32_t &n_full, uint32_t &n_sub, Stage stage)
{
    EvxLaneBlock ref;
    s.bx = px; s.by = py; s.ssd = EVX_BIG; s.sp_index = 0; s.sp_amount = 0; s.sp_enabled = 0;
    evx_block_cost(centre, src, thr, s.sad, s.mad);
    n_full = 1; n_sub = 0;
    if (s.mad < thr) return;                       // already a copy block: no search (motion.cpp:452)
    stage(0, px, py, win);
    // The five rounds and the eight sub-pel directions are ROLLED loops: the search runs on ~28 warps per SM, each at its own
    // place in the code, and what limits them next to other kernels is instruction supply, not the latency a rolled loop
    // adds (unrolled the search is 70 KB of SASS per kernel; -DEVX_K2_UNROLLED keeps that form for A/B runs).
#ifndef EVX_K2_UNROLLED
#pragma unroll 1
#endif
    for (int step = EVX_SEARCH_RADIUS; step > 0; step >>= 1)
    {
        // One 3x3 round (motion.cpp:254-275).  The eight outer cells are costed back to back (no
        // dependence between them, so their loads and arithmetic overlap); the centre is the
        // running best itself, whose sad/mad are already in the state.  Out-of-frame cells still
        // lie inside the zero-filled TMA window, so they are costed too and simply not accepted.
        const int basex = s.bx, basey = s.by;
        // lane c keeps the cost of cell c (lane 4: the centre, i.e. the state's own); the cells two at a time
        int mysad = s.sad, mymad = s.mad;
#pragma unroll CELL_UNROLL
        for (int o = 0; o < 8; ++o)
        {
            const int c = o < 4 ? o : o + 1;
            const int cy = c / 3, cx = c - 3 * cy;
            int csad, cmad;
            evx_load_block(win, basex + (cx - 1) * step, basey + (cy - 1) * step, lane, ref);
            evx_block_cost(ref, src, thr, csad, cmad);
            if (lane == c) { mysad = csad; mymad = cmad; }
        }
        const int lc = lane < 9 ? lane : 0;
        const int x = basex + (lc % 3 - 1) * step, y = basey + (lc / 3 - 1) * step;
        const bool legal = lane < 9 && !(x < 0 || x > g.w - EVX_MB || y < 0 || y > g.h - EVX_MB);
        const int ssd = (x - px) * (x - px) + (y - py) * (y - py);
        const int wl = evx_select_fullpel(s, evx_make_keys(mysad, mymad, ssd, thr, legal), lane, thr, n_full);
        if (wl >= 0)
        {
            s.bx = __shfl_sync(0xFFFFFFFFu, x, wl); s.by = __shfl_sync(0xFFFFFFFFu, y, wl);
            s.sad = __shfl_sync(0xFFFFFFFFu, mysad, wl); s.mad = __shfl_sync(0xFFFFFFFFu, mymad, wl); s.ssd = __shfl_sync(0xFFFFFFFFu, ssd, wl);
        }
        if (step == EVX_SEARCH_RADIUS) stage(1, s.bx, s.by, win);
    }
    // sub-pel (motion.cpp:319-352): eight directions, half and quarter each; lane t takes test t
    EvxLaneBlock best;
    evx_load_block(win, s.bx, s.by, lane, best);
    int tsad = 0, tmad = 0;
#if defined(EVX_K2_UNROLLED) || defined(EVX_K2_UNROLLED_SUBPEL)
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int d8 = 0; d8 < 8; ++d8)
    {
        const int d = d8 < 4 ? d8 : d8 + 1;
        int sh, mh, sq, mq;
        evx_load_block(win, s.bx + d % 3 - 1, s.by + d / 3 - 1, lane, ref);
        // every MAD is computed and the four reductions of a direction are in flight together (thr = -1): the bound
        // sad < 256*thr that spares the full-pel cells their MAD pass costs the sub-pel tests a dependent branch per
        // reduction -- measured 56.5 -> 52.2 us per 1080p frame without it (the full-pel cells measured the other way)
        evx_subpel_cost(best, ref, src, -1, sh, mh, sq, mq);
        tsad = lane == 2 * d8 ? sh : (lane == 2 * d8 + 1 ? sq : tsad);
        tmad = lane == 2 * d8 ? mh : (lane == 2 * d8 + 1 ? mq : tmad);
    }
    {
        const int d8 = (lane >> 1) & 7, d = d8 < 4 ? d8 : d8 + 1;
        const int x = s.bx + d % 3 - 1, y = s.by + d / 3 - 1;
        const bool legal = lane < 16 && !(x < 0 || x > g.w - EVX_MB || y < 0 || y > g.h - EVX_MB);
        const int wt = evx_select_subpel(s, tsad, tmad, legal, lane, thr, n_sub);
        if (wt >= 0)
        {
            const int wd8 = wt >> 1, wd = wd8 < 4 ? wd8 : wd8 + 1;
            s.sp_enabled = 1; s.sp_amount = wt & 1; s.sp_index = evx_frac_index(wd % 3 - 1, wd / 3 - 1);
            s.sad = __shfl_sync(0xFFFFFFFFu, tsad, wt); s.mad = __shfl_sync(0xFFFFFFFFu, tmad, wt);
        }
    }
}

// ------------------------------------------------------------------ K2, one warp per (macroblock, reference)
//
// About four in ten macroblocks of ordinary video are copy blocks whose search ends at the centre test, so the
// unit of work is ONE warp = one (macroblock, reference) search that frees its resources the moment it ends:
//   * the centre test reads the co-located block straight from global memory (no window at all);
//   * a macroblock that does search never needs the whole +-32 range at once.  The step-16 round touches
//     [px-16, px+32) x [py-16, py+32); everything after it (steps 8,4,2,1 and the sub-pel taps) stays within
//     [-16, +32) of that round's winner.  So the window is 48x48 luma + two 24x24 chroma tiles (6.9 KB),
//     fetched by TMA twice (cp.async.bulk.tensor.2d on one mbarrier; out-of-frame samples zero-filled).
// Row pitches of 24 / 12 words keep the lane layout of evx_load_block bank-conflict free
// (rows {0,24,48,72} + 0..7 and {0,12,..,84} + 0..3 tile the 32 banks).
// Two callers: the stand-alone kernel evx_inter_search (grid = (mbw, mbh, refs), one warp per CTA -- the
// kernel the integer roofline is quoted on) and the search follower of the frame pipeline (evx_wavefront.cuh), whose
// warps pull (macroblock, reference) items of one macroblock row and reuse their window and barrier.

#define EVX_K2W_WIN 48
#define EVX_K2W_CWIN 24
#define EVX_K2W_BYTES (EVX_K2W_WIN * EVX_K2W_WIN * 2 + 2 * EVX_K2W_CWIN * EVX_K2W_CWIN * 2)
#define EVX_K2W_SMEM (EVX_K2W_BYTES + 16)
#define EVX_K2W_PER_SM 28         // (6.9 KB + 1 KB reserved) x 28 = 223 KB of the SM's 228 KB

struct EvxK2Params
{
    EvxK2Maps maps;               // 48x48 / 24x24 boxes
    EvxPlanes src, ref[7];
    EvxGeom g;
    EvxInterResult *results;
    unsigned long long *counters;
    int thr;
    int row0;                     // first macroblock row of this launch
};

struct EvxK2Stage
{
    const CUtensorMap *my, *mu, *mv;
    int16_t *wy, *wu, *wv;
    uint64_t *bar;
    uint32_t *phase;              // parity of the barrier's next completion (a persistent warp reuses its barrier)
    int lane;
    __device__ __forceinline__ void fetch(int cx, int cy) const
    {
        if (lane == 0)
        {
            evx_mbar_expect_tx(bar, EVX_K2W_BYTES);
            evx_tma_load_2d(wy, my, cx - 16, cy - 16, bar);
            evx_tma_load_2d(wu, mu, (cx - 16) >> 1, (cy - 16) >> 1, bar);
            evx_tma_load_2d(wv, mv, (cx - 16) >> 1, (cy - 16) >> 1, bar);
        }
    }
    __device__ __forceinline__ void land() const { evx_mbar_wait(bar, *phase & 1u); ++*phase; }
    // k = 0: the window around the macroblock itself was requested before the centre test (it flies while
    // that runs); k = 1: re-centre on the first round's winner -- unless that is the macroblock's own
    // position, whose window is the one already here.
    __device__ __forceinline__ void operator()(int k, int cx, int cy, EvxWin &win) const
    {
        if (k == 0) { land(); return; }
        if (cx - 16 == win.ox && cy - 16 == win.oy) return;
        __syncwarp();                                           // every lane is done reading the previous window
        if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        fetch(cx, cy);
        win.ox = cx - 16; win.oy = cy - 16; win.cox = win.ox >> 1; win.coy = win.oy >> 1;
        land();
    }
};

// One (macroblock, reference) search by one warp.  `win_mem` holds EVX_K2W_BYTES of 128-byte aligned shared memory,
// `bar` an initialised mbarrier (count 1) whose completed phases `phase` counts.  Returns through `out` (lane 0 writes it).
template <int CELL_UNROLL>
__device__ __forceinline__ void evx_k2_item(const CUtensorMap *maps3, const EvxPlanes &srcp, const EvxPlanes &refp, const EvxGeom &g,
                                            int thr, int bx, int by, int ref, int lane, uint8_t *win_mem, uint64_t *bar, uint32_t &phase,
                                            EvxInterResult *out, unsigned long long *counters, uint32_t stamp)
{
    int16_t *wy = reinterpret_cast<int16_t *>(win_mem);
    int16_t *wu = wy + EVX_K2W_WIN * EVX_K2W_WIN;
    int16_t *wv = wu + EVX_K2W_CWIN * EVX_K2W_CWIN;
    const int px = bx * EVX_MB, py = by * EVX_MB;
    EvxK2Stage stage = { maps3, maps3 + 1, maps3 + 2, wy, wu, wv, bar, &phase, lane };
    __syncwarp();                                               // (a persistent warp: the previous item's window reads are over)
    if (lane == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    stage.fetch(px, py);

    EvxLaneSrc src;
    EvxLaneBlock sb, centre;
    evx_load_src_lane(srcp, g, px, py, lane, sb);
    evx_load_src_lane(refp, g, px, py, lane, centre);
    evx_make_src(sb, src);

    EvxWin win;
    win.y = reinterpret_cast<const uint32_t *>(wy); win.u = reinterpret_cast<const uint32_t *>(wu); win.v = reinterpret_cast<const uint32_t *>(wv);
    win.pw_y = EVX_K2W_WIN / 2; win.pw_c = EVX_K2W_CWIN / 2;
    win.ox = px - 16; win.oy = py - 16; win.cox = win.ox >> 1; win.coy = win.oy >> 1;

    EvxSel s;
    uint32_t n_full, n_sub;
    evx_inter_search_warp<CELL_UNROLL>(win, src, centre, g, px, py, thr, lane, s, n_full, n_sub, stage);
    if (n_full == 1) stage.land();      // copy block: the window was never used, but it must have landed before its shared memory is reused or released

    if (lane == 0)
    {
        const EvxDesc d = evx_desc_from_sel(s, 0, ref + 1, px, py, thr);
        *reinterpret_cast<int4 *>(&out->desc) = make_int4((int) d.w0, (int) d.w1, (int) d.w2, (int) d.w3);
        out->sad = s.sad;
        if (stamp) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&out->stamp), "r"(stamp) : "memory");
        else out->stamp = 0u;
        atomicAdd(&counters[0], (unsigned long long) n_full);
        atomicAdd(&counters[1], (unsigned long long) n_sub);
    }
}

// (Measured and rejected: persistent CTAs of eight warps, four per SM = 32 searches per SM, pulling items off a ticket
// counter: 51.7 against 49.2 us per 1080p frame, 128 against 119 us with three references -- the hardware's own CTA
// dispatch of one-warp CTAs is the better ticket machine.)
__global__ void __launch_bounds__(32, EVX_K2W_PER_SM) evx_inter_search(const __grid_constant__ EvxK2Params p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + EVX_K2W_BYTES);
    const int lane = threadIdx.x, ref = blockIdx.z, bx = blockIdx.x, by = blockIdx.y + p.row0;
    if (lane == 0) evx_mbar_init(bar, 1);
    uint32_t phase = 0;
    evx_k2_item<8>(&p.maps.m[ref * 3], p.src, p.ref[ref], p.g, p.thr, bx, by, ref, lane, smem, bar, phase,
                   p.results + (size_t) ref * p.g.mbw * p.g.mbh + (size_t) by * p.g.mbw + bx, p.counters, 0u);
}

// ------------------------------------------------------------------ K3 / K5 shared: transform + reconstruction of one macroblock
//
// Coefficient buffers are block-major: 6 blocks (Y00 Y01 Y10 Y11 U V) x 64.

struct __align__(16) EvxMbShared
{
    int16_t src[384];        // source samples, block-major
    int16_t pred[384];       // prediction, block-major
    int16_t bufa[384];
    int16_t bufb[384];
    int16_t lut[64];         // DCT basis, xftables.h:57-67
    int16_t qmi[64], qmt[64];
    uint32_t recip[128];     // ceil(2^32 / d): exact n/d = umulhi(n, recip[d]) for n < 2^26, 2 <= d < 128
    int red[3 * 32];
    int qp, var;
};

// position of element e (0..383, block-major) inside the macroblock: plane comp, x, y
__device__ __forceinline__ void evx_mb_pos(int e, int &comp, int &x, int &y)
{
    int b = e >> 6, r = (e >> 3) & 7, c = e & 7;
    if (b < 4) { comp = 0; x = (b & 1) * 8 + c; y = (b >> 1) * 8 + r; }
    else { comp = b - 3; x = c; y = r; }
}

// record layout handed to the host: 16x16 luma row-major, then U 8x8, V 8x8
__device__ __forceinline__ int evx_record_index(int e)
{
    int comp, x, y;
    evx_mb_pos(e, comp, x, y);
    return comp == 0 ? y * 16 + x : 256 + (comp - 1) * 64 + y * 8 + x;
}

__device__ __forceinline__ void evx_init_tables(EvxMbShared &sh, int tid, int nt)
{
    for (int k = tid; k < 64; k += nt)
    {
        sh.lut[k] = (int16_t) evx_dct_lut(k >> 3, k & 7);
        sh.qmi[k] = EVX_QM_INTRA[k];
        sh.qmt[k] = EVX_QM_INTER[k];
    }
    for (int d = tid; d < 128; d += nt) sh.recip[d] = d >= 2 ? (uint32_t) ((0x100000000ull + (unsigned) d - 1) / (unsigned) d) : 0u;
}

// rounded_div (math.h:228-236) for 2 <= d < 128 by multiply-high with M = ceil(2^32/d).  With
// M*d = 2^32 + e, 0 <= e < d, floor(a*M / 2^32) == floor(a/d) whenever a*e < 2^32, i.e. for every
// a < 2^25; here a = |n| + d/2 <= 32767*16 + 63 < 2^20.
__device__ __forceinline__ int evx_rdiv_recip(int n, int d, const uint32_t *recip)
{
    uint32_t a = (uint32_t) (n < 0 ? -n : n) + (uint32_t) (d >> 1);
    int q = (int) __umulhi(a, recip[d]);
    return n < 0 ? -q : q;
}

// quantize.cpp:79-180: one coefficient (mode 0 intra luma, 1 intra chroma, 2 inter; pos = j*8+k),
// without hardware division (the divisors are matrix entries 8..45, 2*qp <= 62, dc scales <= 46)
__device__ __forceinline__ int evx_quant_fast(int s, int pos, int mode, int qp, int linear, const int16_t *qm_intra, const int16_t *qm_inter, const uint32_t *recip)
{
    int out;
    if (linear)
    {
        if (mode < 2) out = (short) evx_rdiv_recip(s, qp << 1, recip);
        else { int m = (short) (evx_abs16(s) - (qp >> 1)); out = (short) evx_rdiv_recip(m, qp << 1, recip); out = (short) (out * evx_sign(s)); }
    }
    else if (mode < 2)
    {
        if (pos == 0) out = (short) evx_rdiv_recip(s, mode == 0 ? evx_luma_dc_scale(qp) : evx_chroma_dc_scale(qp), recip);
        else out = (short) evx_rdiv_recip(evx_rdiv_recip(s * 16, qm_intra[pos], recip), qp << 1, recip);
    }
    else
    {
        int f = (short) evx_rdiv_recip(s * 16, qm_inter[pos], recip);
        out = (short) evx_rdiv_recip(f - evx_sign(f) * qp, qp << 1, recip);
    }
    return out;
}

// inverse 8x8 passes (transform.cpp:330-366, 418-433): scale per term
__device__ __forceinline__ int evx_idct_sum(const int16_t *in, int stride, const int16_t *lut, int i)
{
    int t = evx_tdiv_pow2((in[0] * lut[i]) * 45, 7);
#pragma unroll
    for (int k = 1; k < 8; ++k) t += evx_tdiv_pow2(in[k * stride] * lut[k * 8 + i], 1);
    return evx_rdiv_pow2(t, 7);
}

// Reconstruct a non-copy macroblock from quantised coefficients in sh.bufa (block-major):
// dequantise, inverse transform, add the prediction (if any), store into the ring slot.
// (decode.cpp:19-24, 50-72, 107-141)
__device__ __forceinline__ void evx_reconstruct(EvxMbShared &sh, int type, int qp, int linear, bool has_pred,
                                                const EvxPlanes &dst, const EvxGeom &g, int px, int py, int tid, int nt)
{
    bool intra_q = (type & EVX_T_INTRA) && !(type & EVX_T_MOTION);
    for (int e = tid; e < 384; e += nt)
    {
        int b = e >> 6;
        int mode = intra_q ? (b < 4 ? 0 : 1) : 2;
        sh.bufb[e] = (int16_t) evx_dequant(sh.bufa[e], e & 63, mode, qp, linear, sh.qmi, sh.qmt);
    }
    __syncthreads();
    for (int e = tid; e < 384; e += nt)
    {   // vertical pass: scratch[b][i][j] from column j
        int b = e >> 6, i = (e >> 3) & 7, j = e & 7;
        sh.bufa[b * 64 + i * 8 + j] = (int16_t) evx_idct_sum(sh.bufb + b * 64 + j, 8, sh.lut, i);
    }
    __syncthreads();
    int cw = g.w >> 1;
    for (int e = tid; e < 384; e += nt)
    {   // horizontal pass: out[b][j][i] from row j, plus prediction
        int b = e >> 6, j = (e >> 3) & 7, i = e & 7;
        int v = evx_idct_sum(sh.bufa + b * 64 + j * 8, 1, sh.lut, i);
        if (has_pred) v += sh.pred[e];
        int comp, x, y;
        evx_mb_pos(e, comp, x, y);
        if (comp == 0) dst.y[(size_t) (py + y) * g.w + px + x] = (int16_t) v;
        else (comp == 1 ? dst.u : dst.v)[(size_t) ((py >> 1) + y) * cw + (px >> 1) + x] = (int16_t) v;
    }
}

__device__ __forceinline__ void evx_store_pred_as_recon(const EvxMbShared &sh, const EvxPlanes &dst, const EvxGeom &g, int px, int py, int tid, int nt)
{
    int cw = g.w >> 1;
    // (px, py are multiples of 16: 48 aligned 16-byte chunks)
    for (int idx = tid; idx < 48; idx += nt)
    {
        if (idx < 32)
        {
            const int row = idx >> 1, half = idx & 1;
            *reinterpret_cast<uint4 *>(dst.y + (size_t) (py + row) * g.w + px + 8 * half) = *reinterpret_cast<const uint4 *>(sh.pred + ((row >> 3) * 2 + half) * 64 + (row & 7) * 8);
        }
        else
        {
            const int plane = (idx - 32) >> 3, row = (idx - 32) & 7;
            *reinterpret_cast<uint4 *>((plane ? dst.v : dst.u) + (size_t) ((py >> 1) + row) * cw + (px >> 1)) = *reinterpret_cast<const uint4 *>(sh.pred + 256 + plane * 64 + row * 8);
        }
    }
}

// Prediction of a block type straight from a ring slot in global memory (L2 loads: the slot
// may be the frame under construction).  encode.cpp:83-141 / decode.cpp:27-135.
// Full-pel prediction at a position whose x is a multiple of 16 (zero motion above all: most macroblocks of ordinary
// video): chunk idx of the 48 16-byte chunks of a 16x16 + 8x8 + 8x8 block, from the reference planes to a block-major
// buffer (6 blocks x 64, the layout of EvxMbShared::pred).  128-bit loads; idx 0..31 luma (row, half), 32..47 chroma.
__device__ __forceinline__ void evx_pred_chunk16(int16_t *dst, const EvxPlanes &ref, const EvxGeom &g, int bxp, int byp, int idx)
{
    if (idx < 32)
    {
        const int row = idx >> 1, half = idx & 1;
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(ref.y + (size_t) (byp + row) * g.w + bxp + 8 * half));
        *reinterpret_cast<uint4 *>(dst + ((row >> 3) * 2 + half) * 64 + (row & 7) * 8) = v;
    }
    else
    {
        const int plane = (idx - 32) >> 3, row = (idx - 32) & 7;
        const uint4 v = __ldcg(reinterpret_cast<const uint4 *>((plane ? ref.v : ref.u) + (size_t) ((byp >> 1) + row) * (g.w >> 1) + (bxp >> 1)));
        *reinterpret_cast<uint4 *>(dst + 256 + plane * 64 + row * 8) = v;
    }
}

__device__ __forceinline__ void evx_build_pred_global(EvxMbShared &sh, const EvxPlanes &ref, const EvxGeom &g,
                                                      int bxp, int byp, bool sp, int sp_amount, int dx, int dy, int tid, int nt)
{
    int cw = g.w >> 1;
    if (!sp && (bxp & 15) == 0)
    {
        for (int idx = tid; idx < 48; idx += nt) evx_pred_chunk16(sh.pred, ref, g, bxp, byp, idx);
        return;
    }
    for (int e = tid; e < 384; e += nt)
    {
        int comp, x, y;
        evx_mb_pos(e, comp, x, y);
        int a, b = 0;
        if (comp == 0)
        {
            a = __ldcg(ref.y + (size_t) (byp + y) * g.w + bxp + x);
            if (sp) b = __ldcg(ref.y + (size_t) (byp + dy + y) * g.w + bxp + dx + x);
        }
        else
        {
            const int16_t *pl = comp == 1 ? ref.u : ref.v;
            a = __ldcg(pl + (size_t) ((byp >> 1) + y) * cw + (bxp >> 1) + x);
            if (sp) b = __ldcg(pl + (size_t) (((byp + dy) >> 1) + y) * cw + ((bxp + dx) >> 1) + x);
        }
        sh.pred[e] = (int16_t) (sp ? (sp_amount ? evx_lerp_quarter(a, b) : evx_lerp_half(a, b)) : a);
    }
}

// ------------------------------------------------------------------ wavefront scheduling
//
// Macroblock (bx,by) may start when (bx-1,by) and (min(bx+2,W-1),by-1) are complete; this
// satisfies both the read-after-write side of the intra search (reconstructed neighbours)
// and its write-after-read side (stale samples of the ring slot, still needed by the row
// above).  progress[by] = number of completed macroblocks of row `by`.  Work is handed out by
// an atomic ticket in wavefront order, so a CTA holding ticket k only ever waits on tickets
// < k, all of which are already held by running (or finished) CTAs: no deadlock for any grid.

// Polling with an acquire load costs an L1 invalidate (CCTL.IVALL) per poll, which stalls the
// LSU the compute warps of the same SM are using; poll relaxed, fence once on success.
__device__ __forceinline__ unsigned int evx_ld_relaxed_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ int evx_ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Every device-side wait on a counter another CTA advances is BOUNDED.  By construction none can last (tickets are
// claimed in dependency order, and a frame's kernels only wait for kernels launched before them: evx_wavefront.cuh), but a
// wait that did -- a logic error, a foreign context holding the device for seconds -- would otherwise hang the GPU until
// the process is killed.  After `budget_ns` (evxgpu.cu: 4 s, EVXGPU_WAIT_BUDGET_MS) the waiter records what it was
// waiting for in mapped host memory and traps: the launch fails, every later call of the process reports the error.
struct EvxWaitCtx { unsigned long long budget_ns; unsigned int *diag; };

__device__ __forceinline__ unsigned long long evx_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __noinline__ void evx_wait_expired(const EvxWaitCtx &w, unsigned int what, unsigned int a, unsigned int b, unsigned int c)
{
    if (w.diag)
    {
        volatile unsigned int *d = w.diag;
        d[1] = a; d[2] = b; d[3] = c; d[0] = what;
        __threadfence_system();
    }
    __trap();
}

// `poll()` returns true when the wait is over; `ns` is the back-off between polls.  The timer is read every 32nd poll.
#define EVX_BOUNDED_WAIT(w, cond, ns, what, a, b, c)                                                           \
    do {                                                                                                       \
        unsigned long long t0_ = 0; unsigned int n_ = 0;                                                       \
        while (!(cond))                                                                                        \
        {                                                                                                      \
            __nanosleep(ns);                                                                                   \
            if ((++n_ & 31u) == 0u && (w).budget_ns)                                                           \
            {                                                                                                  \
                const unsigned long long t_ = evx_globaltimer();                                               \
                if (!t0_) t0_ = t_;                                                                            \
                else if (t_ - t0_ > (w).budget_ns) evx_wait_expired((w), (what), (a), (b), (c));               \
            }                                                                                                  \
        }                                                                                                      \
    } while (0)

__device__ __forceinline__ void evx_wait_ge(const int *p, int need, const EvxWaitCtx &w)
{
    EVX_BOUNDED_WAIT(w, evx_ld_relaxed(p) >= need, 100, 1u, (unsigned int) need, (unsigned int) evx_ld_relaxed(p), 0u);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// The same for a counter that advances in known steps (a row's progress): far from the target the poll
// backs off (the value cannot arrive sooner than one macroblock time per missing step), next to it the
// poll is tight.
__device__ __forceinline__ void evx_wait_ge_far(const int *p, int need, const EvxWaitCtx &w)
{
    int have;
    EVX_BOUNDED_WAIT(w, (have = evx_ld_relaxed(p)) >= need, (need - have > 2 ? 2000 : (need - have > 1 ? 400 : 40)), 2u, (unsigned int) need, (unsigned int) have, 0u);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

__device__ __forceinline__ void evx_st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------ K3 parameters (kernel in evx_wavefront.cuh)

struct EvxK3Params
{
    EvxPlanes src;             // input_cache
    EvxPlanes ring[8];         // prediction_cache[slot]
    EvxGeom g;
    int R, linear;
    int frame_type, quality;
    uint32_t frame_index;
    EvxInterResult *inter;         // [R-1][nmb], valid when frame_type == 1 (written by this kernel's search role when fuse_k2)
    EvxDesc *table;                // block_table
    int16_t *records;              // [nmb][384] coefficient records, slot = macroblock index
    int *row_records;              // [mbh] non-copy macroblocks per row (for evx_pack_records)
    int *sync;                     // [0] row ticket, [1] total records, then three per-row counters of mbh entries each: progress[] (macroblocks of the
                                   // wavefront row complete), k2c[] (search items of the row claimed), jc[] (deblocking jobs of the tile row claimed); then the row
                                   // tickets of the search follower and of the deblocking follower
    // Frames of one stream pipelined on the device (evxgpu.cu, submit_pipelined): the inter search and the deblocking
    // filter run as ROLES of this kernel, and consecutive frames are gated macroblock by macroblock through per-row
    // counters in device memory.  Cross-frame counters hold frame base + count and are compared cyclically
    // ((int)(value - need) >= 0), so a slot's counters are never zeroed between the frames that reuse it.
    int fuse_k2;                   // 1: the inter search runs beside this kernel (evx_search_follow; results carry `stamp`, the block loaders take what is missing)
    int fuse_dbk;                  // 1: the deblocking filter runs as jobs behind the wavefront (evx_deblock_follow, and the last row's CTA), publishing dbk[]
    int deblocking;                // EVX_ENABLE_DEBLOCKING (with fuse_dbk and no deblocking only the counters advance)
    int thr;                       // (quality >> 2) + 1
    uint32_t stamp;                // this frame's token in EvxInterResult::stamp (non-zero)
    unsigned int *dbk;             // this frame: [mbh] base + tile columns of tile row Y that are filtered (see evx_wavefront.cuh)
    unsigned int dbk_base;
    const unsigned int *prev_dbk;  // the previous frame of the stream (NULL: no predecessor)
    unsigned int prev_base;
    unsigned int *started;         // takes dbk_base when the first CTA of this kernel runs (the next frame is launched behind it)
    EvxWaitCtx wait;
    unsigned long long *counters;
    long long *prof;               // optional [mbh][6] per-row phase cycle sums (NULL in production)
    // for K8 (evx_bins.cuh): what serialize_slice's deltas and DC predictions refer to
    int *prev_motion, *prev_coded; // [nmb] previous macroblock OF THE SAME ROW with a motion vector / with coefficients, or -1
    int *row_last;                 // [2][mbh] last such macroblock of each row, or -1
    EvxK2Maps maps;                // search windows of the inter-search role: [reference][Y,U,V], 48x48 / 24x24 boxes
};

// ------------------------------------------------------------------ K7: pack the non-copy macroblocks' records densely, raster order
// (what serialize_slice consumes, serialize.cpp:125-154).  One CTA per macroblock row.
__global__ void __launch_bounds__(256) evx_pack_records(const EvxDesc *__restrict__ table, const int16_t *__restrict__ records_mb,
                                                        const int *__restrict__ row_records, int16_t *__restrict__ dense, int *total_out, EvxGeom g)
{
    __shared__ int s_idx[256];
    __shared__ int s_wtot[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, by = blockIdx.x;
    if (tid == 0)
    {
        int base = 0, total = 0;
        for (int r = 0; r < g.mbh; ++r) { int c = row_records[r]; if (r < by) base += c; total += c; }
        s_base = base;
        if (by == 0) *total_out = total;
    }
    __syncthreads();
    int running = s_base;
    for (int c0 = 0; c0 < g.mbw; c0 += 256)
    {
        const int m = c0 + tid;
        const bool flag = m < g.mbw && !(table[by * g.mbw + m].w0 & EVX_T_COPY);
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
        if (lane == 0) s_wtot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, ctot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { if (w < warp) woff += s_wtot[w]; ctot += s_wtot[w]; }
        s_idx[tid] = flag ? running + woff + __popc(bal & ((1u << lane) - 1u)) : -1;
        __syncthreads();
        const int cnt = min(256, g.mbw - c0);
        for (int k = warp; k < cnt; k += 8)
        {
            const int idx = s_idx[k];
            if (idx < 0) continue;
            const uint4 *srcp = reinterpret_cast<const uint4 *>(records_mb + (size_t) (by * g.mbw + c0 + k) * 384);
            uint4 *dstp = reinterpret_cast<uint4 *>(dense + (size_t) idx * 384);
            dstp[lane] = srcp[lane];
            if (lane < 16) dstp[32 + lane] = srcp[32 + lane];
        }
        running += ctot;
        __syncthreads();
    }
}

// ------------------------------------------------------------------ K5: decoder reconstruction (decode.cpp:146-170)
//
// The reference decodes macroblocks in raster order.  Only blocks that predict from the frame under
// construction (INTRA_MOTION_*: decode.cpp:27-72) make that order observable:
//   read-after-write   such a block R reads macroblocks X < R (raster) as already decoded this frame;
//   write-after-read   it reads macroblocks X > R as they were BEFORE this frame (stale ring contents),
//                      so X must not be written until R has read them.
// Everything else is independent.  A tiny pre-pass (evx_decode_deps) counts, per macroblock, the
// earlier blocks that still have to read its stale samples; the main kernel hands macroblocks out
// by an atomic ticket in raster order (a block only ever waits on raster-earlier blocks, which are
// already claimed: deadlock-free for any grid), waits on done[] for its decoded sources, signals its
// stale reads, waits for its own readers, then reconstructs.  P-frames -- few intra blocks -- decode
// with the whole GPU in parallel; I-frames serialise only along their real chains.

#define EVX_K5_THREADS 128

struct EvxK5Params
{
    EvxPlanes ring[8];
    EvxGeom g;
    int R, linear;
    uint32_t frame_index;
    const EvxDesc *table;
    const int16_t *records;        // dense, raster order of the non-copy macroblocks
    const int *record_slot;        // [nmb]
    int *sync;                     // [0] ticket
    int *done;                     // [nmb] 1 once the macroblock is reconstructed
    int *readers;                  // [nmb] raster-earlier blocks that have yet to read this block's stale samples
    EvxWaitCtx wait;
};

// source rectangle of a block's prediction, as the macroblocks it touches
struct EvxSrcRect { int bx0, bx1, by0, by1; int bxp, byp, dx, dy; bool sp, current; int slot; };

__device__ __forceinline__ EvxSrcRect evx_decode_source(const EvxDesc &d, const EvxGeom &g, int px, int py, uint32_t frame_index, int R)
{
    EvxSrcRect r;
    const int type = d.type() & 7;
    const int mx = (type & EVX_T_MOTION) ? d.mx() : 0, my = (type & EVX_T_MOTION) ? d.my() : 0;
    r.sp = (type & EVX_T_MOTION) && d.sp_pred();
    r.dx = 0; r.dy = 0;
    if (r.sp) evx_frac_direction(d.sp_index() & 7, r.dx, r.dy);
    const int off = (type & EVX_T_INTRA) ? 0 : d.target();
    r.slot = (int) ((frame_index + (uint32_t) R - (uint32_t) (off % R)) % (uint32_t) R);
    r.current = type != EVX_T_INTRA && r.slot == (int) (frame_index % (uint32_t) R);
    // a valid stream never points outside the frame; clamp so a corrupt one cannot fault
    r.bxp = evx_clip(px + mx, 0, g.w - EVX_MB); r.byp = evx_clip(py + my, 0, g.h - EVX_MB);
    if (r.sp) { r.dx = evx_clip(r.bxp + r.dx, 0, g.w - EVX_MB) - r.bxp; r.dy = evx_clip(r.byp + r.dy, 0, g.h - EVX_MB) - r.byp; }
    const int x0 = min(r.bxp, r.bxp + r.dx), x1 = max(r.bxp, r.bxp + r.dx) + EVX_MB - 1;
    const int y0 = min(r.byp, r.byp + r.dy), y1 = max(r.byp, r.byp + r.dy) + EVX_MB - 1;
    r.bx0 = x0 >> 4; r.bx1 = x1 >> 4; r.by0 = y0 >> 4; r.by1 = y1 >> 4;
    return r;
}

__global__ void __launch_bounds__(256) evx_decode_deps(const EvxDesc *__restrict__ table, EvxGeom g, uint32_t frame_index, int R, int *__restrict__ readers)
{
    const int mb = blockIdx.x * blockDim.x + threadIdx.x;
    if (mb >= g.mbw * g.mbh) return;
    const EvxDesc d = table[mb];
    const int bx = mb % g.mbw, by = mb / g.mbw;
    const EvxSrcRect r = evx_decode_source(d, g, bx * EVX_MB, by * EVX_MB, frame_index, R);
    if (!r.current) return;
    for (int y = r.by0; y <= r.by1; ++y)
    for (int x = r.bx0; x <= r.bx1; ++x)
    {
        const int m = y * g.mbw + x;
        if (m > mb) atomicAdd(&readers[m], 1);
    }
}

__global__ void __launch_bounds__(EVX_K5_THREADS) evx_decode_recon(const __grid_constant__ EvxK5Params p)
{
    __shared__ EvxMbShared sh;
    __shared__ int s_ticket;
    const int tid = threadIdx.x;
    const EvxGeom g = p.g;
    const int nmb = g.mbw * g.mbh;
    const int dest = (int) (p.frame_index % (uint32_t) p.R);
    const EvxPlanes cur = p.ring[dest];

    evx_init_tables(sh, tid, EVX_K5_THREADS);
    for (;;)
    {
        __syncthreads();
        if (tid == 0) s_ticket = atomicAdd(&p.sync[0], 1);
        __syncthreads();
        const int mb = s_ticket;                      // raster order
        if (mb >= nmb) break;
        const int bx = mb % g.mbw, by = mb / g.mbw;
        const int px = bx * EVX_MB, py = by * EVX_MB;
        const EvxDesc d = p.table[mb];
        const int type = d.type() & 7;
        const bool has_pred = type != EVX_T_INTRA;
        const EvxSrcRect r = evx_decode_source(d, g, px, py, p.frame_index, p.R);
        if (!(type & EVX_T_COPY))
        {
            const int16_t *rec = p.records + (size_t) p.record_slot[mb] * 384;
            for (int e = tid; e < 384; e += EVX_K5_THREADS) sh.bufa[e] = rec[evx_record_index(e)];
        }
        if (r.current)
        {   // sources decoded earlier in this frame must be complete
            if (tid == 0)
                for (int y = r.by0; y <= r.by1; ++y)
                for (int x = r.bx0; x <= r.bx1; ++x)
                {
                    const int m = y * g.mbw + x;
                    if (m < mb) evx_wait_ge(p.done + m, 1, p.wait);
                }
            __syncthreads();
        }
        if (has_pred) evx_build_pred_global(sh, p.ring[r.slot], g, r.bxp, r.byp, r.sp, d.sp_amount(), r.dx, r.dy, tid, EVX_K5_THREADS);
        __syncthreads();
        if (tid == 0)
        {
            if (r.current)
            {   // our stale reads are in shared memory now: release the blocks we read from
                __threadfence();
                for (int y = r.by0; y <= r.by1; ++y)
                for (int x = r.bx0; x <= r.bx1; ++x)
                {
                    const int m = y * g.mbw + x;
                    if (m > mb) atomicSub(&p.readers[m], 1);
                }
            }
            // and nobody may still need the samples we are about to overwrite
            EVX_BOUNDED_WAIT(p.wait, evx_ld_relaxed(p.readers + mb) <= 0, 100, 3u, (unsigned int) mb, 0u, 0u);
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        __syncthreads();
        if (type & EVX_T_COPY) evx_store_pred_as_recon(sh, cur, g, px, py, tid, EVX_K5_THREADS);
        else evx_reconstruct(sh, type, d.q_index(), p.linear, has_pred, cur, g, px, py, tid, EVX_K5_THREADS);
        __syncthreads();
        if (tid == 0) { __threadfence(); evx_st_release(p.done + mb, 1); }
    }
}

// ------------------------------------------------------------------ K4: deblocking
//
// The reference filters in place in raster order (deblock.cpp:201-254).  Because every
// filter reads 4+4 and writes at most 3+3 samples on an 8-sample grid, the sweep decomposes
// into INDEPENDENT 8x8 tiles centred on the grid crossings (rows j-4..j+3, cols i-4..i+3):
// inside a tile, the upper band's vertical edge on rows j-4..j-1, then the horizontal edge on
// all eight columns, then the lower band's vertical edge on rows j..j+3; no tile reads or
// writes outside itself (proved against the reference order in tests/test_schedules.py).
// One thread per tile; 16-byte rows -> coalesced 8-byte accesses.

__device__ __forceinline__ void evx_edge_params(const EvxDesc *table, int ia, int ib, int &qp, int &strength)
{
    uint32_t a0 = __ldcg(&table[ia].w0), a3 = __ldcg(&table[ia].w3), b0 = __ldcg(&table[ib].w0), b3 = __ldcg(&table[ib].w3);
    bool ac = (a0 & EVX_T_COPY) != 0, bc = (b0 & EVX_T_COPY) != 0;
    int qa = (a3 >> 8) & 0xFF, qb = (b3 >> 8) & 0xFF;
    qp = (!ac && !bc) ? (qa + qb) >> 1 : (!ac ? qa : (!bc ? qb : 0));        // deblock.cpp:49-65
    strength = (ac && bc) ? 0 : ((ac != bc) ? 1 : 2);                          // deblock.cpp:67-79
}

// deblock.cpp:81-129 on p3..q3 = the eight int16 samples of v (x = p3|p2<<16 ... w = q2|q3<<16).  A CALL, not inlined: a tile
// filters sixteen such lines, and sixteen inlined copies made the tile 50 KB of code -- next to the wavefront rows on the
// same SMs it is instruction supply that limits everybody (evx_wavefront.cuh, evx_k3_compute).
__device__ __noinline__ uint4 evx_filter8(uint4 v, int qp, int strength, int luma)
{
    const int p3 = evx_lo16(v.x), p2 = evx_hi16(v.x), p1 = evx_lo16(v.y), p0 = evx_hi16(v.y);
    const int q0 = evx_lo16(v.z), q1 = evx_hi16(v.z), q2 = evx_lo16(v.w), q3 = evx_hi16(v.w);
    const int d0 = (short) abs(p0 - q0), d1 = (short) abs(p1 - p0), d2 = (short) abs(q1 - q0);
    const int al = EVX_ALPHA[qp & 31], be = EVX_BETA[qp & 31];
    if (d0 >= al || d1 >= be || d2 >= be) return v;
    int s1 = p2, s2 = p1, s3 = p0, s4 = q0, s5 = q1, s6 = q2;
    if (strength == 2)
    {
        s3 = evx_rdiv_pow2(p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1, 3);
        s2 = evx_rdiv_pow2(p2 + p1 + p0 + q0, 2);
        s4 = evx_rdiv_pow2(p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2, 3);
        s5 = evx_rdiv_pow2(p0 + q0 + q1 + q2, 2);
        if (luma)
        {
            s1 = evx_rdiv_pow2(2 * p3 + 3 * p2 + p1 + p0 + q0, 3);
            s6 = evx_rdiv_pow2(2 * q3 + 3 * q2 + q1 + q0 + p0, 3);
        }
    }
    else if (strength == 1)
    {
        s3 = evx_rdiv_pow2(((q0 + p0) * 4) + p1 - q1, 3);
        s4 = evx_rdiv_pow2(((q0 + p0) * 4) + q1 - p1, 3);
        if (luma)
        {
            s2 = evx_rdiv_pow2((p2 * 4) + (p0 * 2) + (q0 * 2), 3);
            s5 = evx_rdiv_pow2((q2 * 4) + (q0 * 2) + (p0 * 2), 3);
        }
    }
    // (every result is stored as a short, deblock.cpp: the pack keeps the low 16 bits)
    return make_uint4(evx_pack16(p3, s1), evx_pack16(s2, s3), evx_pack16(s4, s5), evx_pack16(s6, q3));
}

struct EvxK4Params
{
    EvxPlanes pl; EvxGeom g; const EvxDesc *table;
};

// One 8x8 tile of plane `comp` centred on the grid crossing (i, j) = (8 tx, 8 ty), 0 <= tx <= w/8, 0 <= ty <= h/8.
__device__ __forceinline__ void evx_deblock_tile(const EvxPlanes &pl, const EvxGeom &g, const EvxDesc *table, int comp, int tx, int ty)
{
    const bool luma = comp == 0;
    const int w = luma ? g.w : g.w >> 1, h = luma ? g.h : g.h >> 1;
    const int mbs = luma ? 16 : 8;
    const int wb = w / mbs;
    int16_t *img = comp == 0 ? pl.y : (comp == 1 ? pl.u : pl.v);
    const int i = tx * 8, j = ty * 8;
    const bool has_l = i > 0, has_r = i < w, has_t = j > 0, has_b = j < h;

    // the tile as eight packed rows: t[r] = columns i-4..i+3 of row j-4+r, two int16 samples per word
    uint4 t[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        bool rv = r < 4 ? has_t : has_b;
        uint2 a = make_uint2(0, 0), b = make_uint2(0, 0);
        if (rv && has_l) a = __ldcg(reinterpret_cast<const uint2 *>(img + (size_t) (j - 4 + r) * w + i - 4));
        if (rv && has_r) b = __ldcg(reinterpret_cast<const uint2 *>(img + (size_t) (j - 4 + r) * w + i));
        t[r] = make_uint4(a.x, a.y, b.x, b.y);
    }
    int qp, st;
    // 1. vertical edge at column i, upper band (rows j-4..j-1 belong to band j-8)
    if (has_l && has_r && has_t)
    {
        int brow = ((j - 8) / mbs) * wb;
        evx_edge_params(table, (i - 1) / mbs + brow, i / mbs + brow, qp, st);
        if (st)
        {
#pragma unroll
            for (int r = 0; r < 4; ++r) t[r] = evx_filter8(t[r], qp, st, luma);
        }
    }
    // 2. horizontal edge at row j: columns i-4..i-1 belong to edge segment i-8, columns i..i+3 to segment i.  A column
    // pair (2m, 2m+1) is word m of every row: the two columns are gathered by byte permutes, filtered, scattered back.
    if (has_t && has_b)
    {
#pragma unroll
        for (int half = 0; half < 2; ++half)
        {
            if (half == 0 ? !has_l : !has_r) continue;
            int col = half == 0 ? i - 8 : i;
            evx_edge_params(table, col / mbs + ((j - 1) / mbs) * wb, col / mbs + (j / mbs) * wb, qp, st);
            if (!st) continue;
#pragma unroll
            for (int m = 2 * half; m < 2 * half + 2; ++m)
            {
                uint32_t wr[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) wr[r] = m == 0 ? t[r].x : m == 1 ? t[r].y : m == 2 ? t[r].z : t[r].w;
                uint4 c0 = make_uint4(__byte_perm(wr[0], wr[1], 0x5410), __byte_perm(wr[2], wr[3], 0x5410), __byte_perm(wr[4], wr[5], 0x5410), __byte_perm(wr[6], wr[7], 0x5410));
                uint4 c1 = make_uint4(__byte_perm(wr[0], wr[1], 0x7632), __byte_perm(wr[2], wr[3], 0x7632), __byte_perm(wr[4], wr[5], 0x7632), __byte_perm(wr[6], wr[7], 0x7632));
                c0 = evx_filter8(c0, qp, st, luma);
                c1 = evx_filter8(c1, qp, st, luma);
                const uint32_t a0[4] = { c0.x, c0.y, c0.z, c0.w }, a1[4] = { c1.x, c1.y, c1.z, c1.w };
#pragma unroll
                for (int r = 0; r < 8; ++r)
                {
                    const uint32_t v = (r & 1) ? __byte_perm(a0[r >> 1], a1[r >> 1], 0x7632) : __byte_perm(a0[r >> 1], a1[r >> 1], 0x5410);
                    if (m == 0) t[r].x = v; else if (m == 1) t[r].y = v; else if (m == 2) t[r].z = v; else t[r].w = v;
                }
            }
        }
    }
    // 3. vertical edge at column i, lower band (rows j..j+3 belong to band j)
    if (has_l && has_r && has_b)
    {
        int brow = (j / mbs) * wb;
        evx_edge_params(table, (i - 1) / mbs + brow, i / mbs + brow, qp, st);
        if (st)
        {
#pragma unroll
            for (int r = 4; r < 8; ++r) t[r] = evx_filter8(t[r], qp, st, luma);
        }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r)
    {
        bool rv = r < 4 ? has_t : has_b;
        if (rv && has_l) *reinterpret_cast<uint2 *>(img + (size_t) (j - 4 + r) * w + i - 4) = make_uint2(t[r].x, t[r].y);
        if (rv && has_r) *reinterpret_cast<uint2 *>(img + (size_t) (j - 4 + r) * w + i) = make_uint2(t[r].z, t[r].w);
    }
}

// the whole frame, one thread per tile: blockIdx.z = plane
__global__ void __launch_bounds__(128) evx_deblock(const __grid_constant__ EvxK4Params p)
{
    const int comp = blockIdx.z;
    const int w = comp == 0 ? p.g.w : p.g.w >> 1, h = comp == 0 ? p.g.h : p.g.h >> 1;
    const int tx = blockIdx.x * blockDim.x + threadIdx.x, ty = (int) blockIdx.y;
    if (tx > w / 8 || ty > h / 8) return;
    evx_deblock_tile(p.pl, p.g, p.table, comp, tx, ty);
}
