// pb_tiled2.cuh -- the two-lens / blend-band tile class of a batch through a double-fisheye source
// (remap_tiled_kernel CLS = 2) with 512 threads per CTA.
//
// That class stages two rectangles of ~17 KB per frame, so its CTAs need ~100 KB of shared memory
// and only two fit an SM.  With 256 threads that is 16 warps per SM, and the experiment build that
// switches the loads off (profiles/experiments/README.md, "cfg5 x16 by parts") shows the class bound
// by its own gather / blend instruction stream at that occupancy: 0.272 of its 0.327 ms with every
// load removed, issue slots 44 % busy.  The shared memory cannot shrink, but the threads can
// double: here a thread owns ONE quad (4 consecutive pixels of one row) instead of two, a CTA is 16
// warps, an SM holds 32 -- at ~60 registers per thread instead of 124.  Everything else is the lean
// frame loop of pb_tiled.cuh: same tables, same float64 operations in the same order (bit-identical
// offsets), same ring of frame groups filled by TMA box loads, same guarded fixed-point blend, same
// TMA tile store.
#pragma once

#include "pb_tiled.cuh"

namespace pb {

constexpr int kTile2Threads = 512;

// separable tables (un-rotated equirect output), double-fisheye source, C = 3; tiles from a.tile_list
__global__ void __launch_bounds__(kTile2Threads, 2) remap_two_lens_kernel(const __grid_constant__ TiledArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    // [ out tiles: 2 x 6144 ][ stage area: stage_bytes + 128 ][ TileShared ]
    unsigned char* out_tiles = smem;
    unsigned char* stages = smem + 2 * kOutTileBytes;
    TileShared* sh = reinterpret_cast<TileShared*>(stages + a.stage_bytes + 128);

    const int tid = threadIdx.x;
    const int qc = tid & (kQuadsPerRow - 1);
    const int rg = tid >> 3;  // 0..63: the row of this thread's quad
    int tile_x, tile_y;
    {
        const int t = a.tile_list ? __ldg(a.tile_list + blockIdx.x) : (int)blockIdx.x;
        tile_y = t / a.tiles_x;
        tile_x = t - tile_y * a.tiles_x;
    }
    tile_y += a.tile_y0;
    const int x0 = tile_x * kTileW;
    const int y0 = tile_y * kTileH;
    const int ys = y0 - a.tile_y0 * kTileH;

    if (tid == 0) {
        ptx::prefetch_tensormap(&a.dst_map);
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) ptx::mbarrier_init(&sh->bar[g], 1);
        ptx::fence_mbarrier_init();
    }

    // ------------------------------------------------------------ tables, footprints
    int4 fpv[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) fpv[s] = __ldg(a.tile_fp + (tile_y * a.tiles_x + tile_x) * 2 + s);
    double2 cs[4];
    {
        const double2* __restrict__ col = reinterpret_cast<const double2*>(a.sep1_col) + tile_x * kTileW;
#pragma unroll
        for (int k = 0; k < 4; ++k) cs[k] = __ldg(col + k * 8 + qc);
    }
    const double2* __restrict__ row = reinterpret_cast<const double2*>(a.sep1_row) + 2 * (tile_y * kTileH + rg);
    const double2 r01 = __ldg(row);      // lens radius of this row: left, right
    const double2 wrow = __ldg(row + 1);  // blend weights of this row: left, right
    const bool unit_weights = __syncthreads_and(wrow.x == 1.0 && wrow.y == 1.0);  // (also: barriers initialised)

    int by0[2], xb0[2], nbox[2], pitch[2];
    bool all_valid[2];
    bool wide = false;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        by0[s] = fpv[s].x;
        xb0[s] = fpv[s].y;
        nbox[s] = fpv[s].z;
        all_valid[s] = fpv[s].w & 1;
        const int units = stage_units(fpv[s].w >> 1);
        pitch[s] = 16 * units;
        if (nbox[s] > 0 && units > a.max_units) wide = true;
    }
    const int rect0 = nbox[0] * kBoxRows * pitch[0];
    const int rect1 = nbox[1] * kBoxRows * pitch[1];
    const int group_bytes = 128 + rect0 + rect1;
    const int n_groups = min(min(kMaxGroups, a.n_frames), (a.stage_bytes + 128) / group_bytes);
    if (wide || n_groups < 1) {  // block-uniform: a footprint that cannot be staged (a pole)
        direct_tile<PB_KIND_EQUIRECT, PB_KIND_DOUBLE, 1>(a, nullptr, nullptr, out_tiles, x0, y0);
        return;
    }
    auto issue_group = [&](int f, int g) {  // one thread: both rectangles of frame f into group g
        unsigned char* base = stages + g * group_bytes + 128;
        ptx::mbarrier_arrive_expect_tx(&sh->bar[g], (unsigned)(rect0 + rect1));
        const uint64_t keep = ptx::policy_of(a.load_policy);
        if (nbox[0] > 0) {
            const CUtensorMap* map = &a.src_maps[(pitch[0] >> 5) - (kMinStageUnits >> 1)];
            for (int k = 0; k < nbox[0]; ++k)
                ptx::tma_load_3d_hint(base + k * kBoxRows * pitch[0], map, xb0[0] >> 1, by0[0] + k * kBoxRows, f, &sh->bar[g], keep);
        }
        if (nbox[1] > 0) {
            const CUtensorMap* map = &a.src_maps[(pitch[1] >> 5) - (kMinStageUnits >> 1)];
            for (int k = 0; k < nbox[1]; ++k)
                ptx::tma_load_3d_hint(base + rect0 + k * kBoxRows * pitch[1], map, xb0[1] >> 1, by0[1] + k * kBoxRows, f,
                                      &sh->bar[g], keep);
        }
    };
    if (tid < n_groups * 8) reinterpret_cast<int4*>(stages + (tid >> 3) * group_bytes)[tid & 7] = make_int4(0, 0, 0, 0);
    ptx::fence_async_smem();
    if (tid == 0) {
        for (int f = 0; f < n_groups; ++f) issue_group(f, f);
    }

    // ------------------------------------------------------------ resolve: the address of every pixel inside a frame group
    const unsigned stages_sa = ptx::smem_addr(stages);
    unsigned adr[2][4], shf[2][4];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const bool right = s != 0;
        const int w = right ? a.src.wr : a.src.wl;
        const double cx = right ? a.src.cxr : a.src.cxl;
        const int origin = by0[s] * pitch[s] + xb0[s];
        const double dist = s ? r01.y : r01.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned rel = 0;  // the group's zero bytes: no source pixel
            if (nbox[s] > 0) {
                double fx, fy;
                camera_fxy(cs[k].x, cs[k].y, dist, a.src.cy, cx, fx, fy);
                int px = trunc_abs(fx);
                if (right) px = a.src.W - 1 - px;
                const int off = trunc_abs(fy) * pitch[s] + (px * 3 - origin);
                if (all_valid[s] || inside_image(fx, fy, w, a.src.H)) rel = (unsigned)(off + 128 + (s ? rect0 : 0));
            }
            adr[s][k] = stages_sa + (rel & ~3u);
            shf[s][k] = rel << 3;
        }
    }
    // blend of this thread's row: 0 = unit weights (exact byte add), 1 = guarded fixed point, 2 = float64
    unsigned wfix0 = 0, wfix1 = 0;
    int wmode = 0;
    if (!unit_weights && !(wrow.x == 1.0 && wrow.y == 1.0)) wmode = fix_weights(wrow.x, wrow.y, wfix0, wfix1) ? 1 : 2;
    __syncthreads();  // the groups' zero bytes are in place

    const uint64_t drop = ptx::policy_of(a.store_policy);
    const unsigned out_sa = ptx::smem_addr(out_tiles) + rg * kOutRowBytes + qc * 12;
    const unsigned bar_sa = ptx::smem_addr(&sh->bar[0]);
    auto frame_loop = [&](auto ACT, auto WGT) {
        constexpr int act = decltype(ACT)::value;   // 1: left lens only, 2: right lens only, 3: both
        constexpr bool wgt = decltype(WGT)::value;  // some row of the tile has a weighted blend
        int g = 0;
        unsigned parity = 0;
        for (int f = 0; f < a.n_frames; ++f) {
            const unsigned goff = (unsigned)(g * group_bytes);
            ptx::mbarrier_wait_sa(bar_sa + 8 * g, parity);
            unsigned v[4], w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k] = (act & 1) ? ptx::lds_pixel(adr[0][k], goff, shf[0][k]) : 0u;
                w[k] = (act & 2) ? ptx::lds_pixel(adr[1][k], goff, shf[1][k]) : 0u;
            }
            if (!wgt || wmode == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (act == 3) ? __vadd4(v[k], w[k]) : (act == 1 ? v[k] : w[k]);
            } else if (wmode == 1) {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = blend_px_fix(v[k], wfix0, wrow.x, w[k], wfix1, wrow.y);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = blend_px_weighted(v[k], wrow.x, w[k], wrow.y);
            }
            store_quad_sa(out_sa + (f & 1) * kOutTileBytes, v);
            ptx::fence_async_smem();
            if (tid == 0) ptx::bulk_wait_read0();  // stores up to frame f - 1 have left their tiles
            __syncthreads();
            if (tid == 0) {
                if (f + n_groups < a.n_frames) issue_group(f + n_groups, g);
                ptx::tma_store_3d_hint(&a.dst_map, x0 * 3, ys, f, out_tiles + (f & 1) * kOutTileBytes, drop);
                ptx::bulk_commit();
            }
            if (++g == n_groups) {
                g = 0;
                parity ^= 1u;
            }
        }
    };
    const int n_act = (nbox[0] > 0) + (nbox[1] > 0);
    if (!unit_weights) {
        if (n_act == 2) frame_loop(std::integral_constant<int, 3>{}, std::true_type{});
        else if (nbox[1] == 0) frame_loop(std::integral_constant<int, 1>{}, std::true_type{});
        else frame_loop(std::integral_constant<int, 2>{}, std::true_type{});
    } else if (n_act == 2) frame_loop(std::integral_constant<int, 3>{}, std::false_type{});
    else if (nbox[1] == 0) frame_loop(std::integral_constant<int, 1>{}, std::false_type{});
    else frame_loop(std::integral_constant<int, 2>{}, std::false_type{});
    if (tid == 0) ptx::bulk_wait_read0();
}

inline int two_lens_smem_bytes(int stage_bytes) {
    return 2 * kOutTileBytes + stage_bytes + 128 + (int)sizeof(TileShared) + 128;
}

}  // namespace pb
