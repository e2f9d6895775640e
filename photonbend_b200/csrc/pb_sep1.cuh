// pb_sep1.cuh -- ONE frame through a separable geometry (un-rotated equirect output from a camera
// or double-fisheye source): what a make-pano call does (BASELINE configs 1 and 4, the 8K target).
//
// The batched kernel (pb_tiled.cuh) resolves a tile's source offsets once and applies them to
// every frame; with a single frame that set-up is all there is, and ncu shows the launch bound by
// instruction issue (59 thread-instructions per pixel around 6 FP64 operations) and by the
// load -> gather -> store latency chain of one-tile CTAs (profiles/r1_ncu_full_T_1frame.txt).
// This kernel does the same work with
//   * persistent CTAs that walk tiles u = blockIdx.x, + gridDim.x, ... of a per-plan table in
//     launch order (tile position and footprint pre-decoded: no raster divisions, no footprint
//     arithmetic in the kernel),
//   * a two-deep pipeline per CTA: the TMA box loads of tile n+1 are in flight, and those of tile
//     n+2 are issued, while tile n is gathered; its TMA tile store drains under tile n+1,
//   * per pixel: 4 FP64 multiply/adds + 2 FP64 truncating adds, two integer multiply-adds, two
//     LDS.32 and a funnel shift -- offsets are consumed as they are produced (no offset arrays),
//   * one block barrier per tile.
// Arithmetic per pixel is exactly that of pb_tiled.cuh's separable mode (same tables, same
// operations in the same order), so results are bit-identical by construction.
#pragma once

#include "pb_tiled.cuh"

namespace pb {

// Per (tile in launch order, slot) descriptor, written once per plan by pb_sep1_table_kernel:
//   x: first source row (bits 0-14) | number of 16-row boxes << 16 | staged-row units << 24 | all_valid << 30
//   y: first staged byte of a source row (multiple of 16)
//   z: tile_x | tile_y << 16
//   w: by0 * pitch + xb0  (byte offset of the rectangle's origin in "staged pitch" coordinates)
// With a tile list (a class of tiles of a double-fisheye source, pb_plan): entry u describes tile
// list[u]; one_lens: ONE descriptor per tile, that of the lens the tile sees, bit 31 of x = right lens.
__global__ void __launch_bounds__(256) pb_sep1_table_kernel(const int4* __restrict__ tile_fp, int4* __restrict__ tab,
                                                            int tiles_x, int tiles_y, int raster_band, int nslot,
                                                            const int* __restrict__ list = nullptr, int n_list = 0,
                                                            int one_lens = 0) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (list ? n_list : tiles_x * tiles_y)) return;
    int tile_x, tile_y;
    if (list) {
        tile_y = list[u] / tiles_x;
        tile_x = list[u] - tile_y * tiles_x;
    } else if (raster_band > 0) {
        const int per_band = raster_band * tiles_x;
        const int band = u / per_band, within = u - band * per_band;
        const int bh = min(raster_band, tiles_y - band * raster_band);
        tile_x = within / bh;
        tile_y = band * raster_band + (within - tile_x * bh);
    } else {
        tile_y = u / tiles_x;
        tile_x = u - tile_y * tiles_x;
    }
    for (int s = 0; s < nslot; ++s) {
        const int4 fp = tile_fp[(tile_y * tiles_x + tile_x) * nslot + s];
        int4 d = make_int4(0, 0, tile_x | (tile_y << 16), 0);
        if (fp.z > 0) {
            const int units = stage_units(fp.w >> 1);
            const bool fits = units <= kMaxStageUnits && fp.z <= 255;
            // a footprint that cannot be staged keeps nbox = 255 / units = 63: the kernel then gathers from global memory
            d.x = fp.x | ((fits ? fp.z : 255) << 16) | ((fits ? units : 63) << 24) | ((fp.w & 1) << 30);
            d.y = fp.y;
            d.w = fp.x * 16 * units + fp.y;
        }
        if (!one_lens) tab[u * nslot + s] = d;
        else if (fp.z > 0) tab[u] = make_int4(d.x | (s ? (int)0x80000000u : 0), d.y, d.z, d.w);
    }
}

// Per-tile table slices for the single-frame kernel, laid out so that ONE bulk copy brings a tile's
// slice into shared memory and the threads' reads of it are conflict-free:
//   col: tiles_x x 32 entries (cos, sin); column c of a tile sits at (c & 3) * 8 + (c >> 2), so the
//        eight quad columns of a warp read 128 contiguous bytes for each of their 4 columns;
//   row: tiles_y x 64 entries; camera: lens radius (8 bytes); double: (radius left, radius right,
//        weight left, weight right) (32 bytes).
// Entries past the image edge repeat the last column / row (their pixels are clipped by the store).
__global__ void __launch_bounds__(256) pb_sep1_slices_kernel(const double* __restrict__ col_tab,
                                                             const double* __restrict__ row_tab, double* __restrict__ col_out,
                                                             double* __restrict__ row_out, int W, int H, int tiles_x,
                                                             int tiles_y, int row_doubles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < tiles_x * kTileW) {
        const int c = t & (kTileW - 1);
        const int j = min(t, W - 1);
        double* o = col_out + 2 * ((t - c) + (c & 3) * 8 + (c >> 2));
        o[0] = col_tab[2 * j];
        o[1] = col_tab[2 * j + 1];
    }
    if (t < tiles_y * kTileH) {
        const int i = min(t, H - 1);
        for (int e = 0; e < row_doubles; ++e) row_out[(size_t)t * row_doubles + e] = row_tab[4 * i + e];
    }
}

// The 3 bytes of the staged pixel at byte offset off.  Single-lens source: the second word is
// loaded only by the lanes that need it (T x1 47.0 -> 45.4 us); double source: both words always
// (the predicate costs more than the conflicts it saves there: 74.4 vs 76.2 us).
#ifndef PB_SEP1_PRED
#define PB_SEP1_PRED 1
#endif
template <bool PRED>
__device__ __forceinline__ unsigned sep1_pick(unsigned stage_sa, int off) {
    if (PRED && PB_SEP1_PRED) return ptx::lds_pixel_pred(stage_sa + (unsigned)off);
    return ptx::lds_pixel(stage_sa, (unsigned)off & ~3u, (unsigned)off << 3);
}

struct Sep1Slot {
    int nbox, pitch, rect, by0, xb0, origin, tile;
    bool all_valid;
    __device__ __forceinline__ void decode(const int4 d) {
        tile = d.z;
        nbox = (d.x >> 16) & 0xff;
        pitch = ((d.x >> 24) & 0x3f) << 4;
        rect = nbox * kBoxRows * pitch;
        by0 = d.x & 0x7fff;
        xb0 = d.y;
        origin = d.w;
        all_valid = (d.x >> 30) & 1;
    }
};

#ifndef PB_SEP1_CAM_CTAS
#define PB_SEP1_CAM_CTAS 4  // 60 registers, no spills: T x1 43.2 us (5 CTAs at 48 registers spill: 45.4 us)
#endif
// NB: stage buffers = depth of the load pipeline (tile n is gathered while the loads of tiles
// n+1 .. n+NB-1 are in flight and those of tile n+NB are issued)
// CLS (double-fisheye source; the plan sorts the tiles into two launches like the batched kernel):
// 0 = every tile, 1 = the tiles that see exactly one lens at unit weights (one slot, chosen per
// tile; small buffers, 4 CTAs per SM), 2 = the rest (the code of class 0 over a tile list)
template <int SRC_KIND, int NB, int CLS = 0>
__global__ void __launch_bounds__(kTileThreads, (SRC_KIND == PB_KIND_DOUBLE && CLS != 1) ? 3 : PB_SEP1_CAM_CTAS)
remap_sep1_kernel(const __grid_constant__ TiledArgs a) {
    static_assert(NB >= 2 && NB <= 4, "2..4 stage buffers");
    constexpr bool DBL = (SRC_KIND == PB_KIND_DOUBLE);
    constexpr bool ONE = DBL && CLS == 1;
    constexpr int NSLOT = (DBL && !ONE) ? 2 : 1;

    extern __shared__ __align__(128) unsigned char smem[];
    // [ out tile 0 ][ out tile 1 ][ NB stage buffers ][ ring of 8 tile descriptors ][ NB mbarriers ]
    // stage buffer: [128 zero bytes][column slice 512][row slice 512 | 2048][source rectangles, up to cap bytes]
    constexpr int kColBytes = kTileW * 16;
    constexpr int kRowBytes = kTileH * (DBL ? 32 : 8);
    constexpr int kHead = 128 + kColBytes + kRowBytes;  // rectangles start here (a multiple of 128)
    const int cap = a.sep1_cap;
    const int buf_bytes = kHead + cap;
    unsigned char* out_tiles = smem;
    unsigned char* stages = smem + 2 * kOutTileBytes;
    int4* ring = reinterpret_cast<int4*>(stages + NB * buf_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 16);

    const int tid = threadIdx.x;
    const int qc = tid & (kQuadsPerRow - 1);
    const int rg = tid >> 3;
    const int G = gridDim.x;
    const int n_tiles = a.tile_list ? a.n_list : a.tiles_x * a.tiles_y;  // (a list: sep1_tab is that class's table)
    const int4* __restrict__ tab = a.sep1_tab;

    if (tid == 0) {
        ptx::prefetch_tensormap(&a.dst_map);
#pragma unroll
        for (int b = 0; b < NB; ++b) ptx::mbarrier_init(&bars[b], 1);
        ptx::fence_mbarrier_init();
    }
    if (tid < 8 * NB) reinterpret_cast<int4*>(stages + (tid >> 3) * buf_bytes)[tid & 7] = make_int4(0, 0, 0, 0);
    if (tid < (NB + 1) * NSLOT) {  // descriptors of this CTA's first NB + 1 tiles
        const int k = tid / NSLOT, s = tid - k * NSLOT;
        const int u = blockIdx.x + k * G;
        ring[k * 2 + s] = (u < n_tiles) ? __ldg(tab + u * NSLOT + s) : make_int4(0, 0, 0, 0);
    }
    // Programmatic dependent launch: the next grid on this stream may be set up while this one
    // runs, and everything above -- barriers, zero bytes, descriptors out of the plan's tables,
    // which no kernel writes after the plan was made -- ran before the grids ahead of this one had
    // finished.  From here on the source image and the output buffer are touched: wait for them.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // one thread: the box loads of the tile whose descriptors sit in ring slot r, into stage buffer b
    auto issue = [&](int r, int b) {
        Sep1Slot s0, s1;
        s0.decode(ring[r * 2]);
        s1.decode(NSLOT == 2 ? ring[r * 2 + 1] : make_int4(0, 0, 0, 0));
        const int tx = s0.tile & 0xffff, ty = s0.tile >> 16;
        int total = s0.rect + s1.rect;
        // gathered from global memory (only the table slices are staged): too large for a buffer, or
        // wider than the widest box this launch has a tensor map for
        if (total > cap || max(s0.pitch, s1.pitch) > 16 * a.max_units) total = 0;
        ptx::mbarrier_arrive_expect_tx(&bars[b], (unsigned)(total + kColBytes + kRowBytes));
        unsigned char* head = stages + b * buf_bytes + 128;
        ptx::bulk_g2s(head, a.sep1_col + (size_t)tx * (kColBytes / 8), kColBytes, &bars[b]);
        ptx::bulk_g2s(head + kColBytes, a.sep1_row + (size_t)ty * (kRowBytes / 8), kRowBytes, &bars[b]);
        if (total == 0) return;
        const uint64_t keep = ptx::policy_evict_last();
        unsigned char* base = stages + b * buf_bytes + kHead;
        if (s0.nbox > 0) {
            const CUtensorMap* map = &a.src_maps[(s0.pitch >> 5) - (kMinStageUnits >> 1)];
            for (int k = 0; k < s0.nbox; ++k)
                ptx::tma_load_3d_hint(base + k * kBoxRows * s0.pitch, map, s0.xb0 >> 1, s0.by0 + k * kBoxRows, 0, &bars[b], keep);
        }
        if (NSLOT == 2 && s1.nbox > 0) {
            const CUtensorMap* map = &a.src_maps[(s1.pitch >> 5) - (kMinStageUnits >> 1)];
            for (int k = 0; k < s1.nbox; ++k)
                ptx::tma_load_3d_hint(base + s0.rect + k * kBoxRows * s1.pitch, map, s1.xb0 >> 1, s1.by0 + k * kBoxRows, 0,
                                      &bars[b], keep);
        }
    };
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NB; ++b)
            if ((int)blockIdx.x + b * G < n_tiles) issue(b, b);
    }

    const uint64_t drop = ptx::policy_evict_first();
    const unsigned stages_sa = ptx::smem_addr(stages);
    const unsigned out_sa = ptx::smem_addr(out_tiles) + rg * kOutRowBytes + qc * 12;
    const double cy = a.src.cy;
    unsigned phase = 0;  // bit b: parity the next wait on stage buffer b expects

    int it = 0, b = 0;
    for (int u = blockIdx.x; u < n_tiles; u += G, ++it, b = (b + 1 == NB) ? 0 : b + 1) {
        const int ob = it & 1;  // output tile
        const int4 d0 = ring[(it & 7) * 2];
        const int4 d1 = NSLOT == 2 ? ring[(it & 7) * 2 + 1] : make_int4(0, 0, 0, 0);
        const bool right_lens = ONE && d0.x < 0;  // ONE: the lens this tile sees
        // the descriptor of the tile NB + 1 ahead travels while this tile is processed
        int4 pre = make_int4(0, 0, 0, 0);
        if (tid < NSLOT && u + (NB + 1) * G < n_tiles) pre = __ldg(tab + (u + (NB + 1) * G) * NSLOT + tid);

        Sep1Slot sl[2];
        sl[0].decode(d0);
        sl[1].decode(d1);
        const int x0 = (d0.z & 0xffff) * kTileW, y0 = (d0.z >> 16) * kTileH;
        const int total = sl[0].rect + sl[1].rect;
        const bool staged = total <= cap && max(sl[0].pitch, sl[1].pitch) <= 16 * a.max_units;  // block-uniform

        ptx::mbarrier_wait_sa(ptx::smem_addr(&bars[b]), (phase >> b) & 1u);
        phase ^= 1u << b;
        const unsigned stage_sa = stages_sa + b * buf_bytes;
        const unsigned char* __restrict__ frame = a.src_px;

        // table slices: 4 columns, 2 rows per thread
        const unsigned char* head = stages + b * buf_bytes + 128;
        double2 cs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) cs[k] = reinterpret_cast<const double2*>(head)[k * 8 + qc];
        double2 r01[kRowsPerThread], r23[kRowsPerThread];
#pragma unroll
        for (int q = 0; q < kRowsPerThread; ++q) {
            const int r = rg + q * kRowGroups;
            if (ONE) {
                r01[q].x = reinterpret_cast<const double*>(head + kColBytes)[4 * r + (right_lens ? 1 : 0)];
                r01[q].y = 0.0;
                r23[q] = make_double2(1.0, 1.0);
            } else if (DBL) {
                r01[q] = reinterpret_cast<const double2*>(head + kColBytes)[2 * r];
                r23[q] = reinterpret_cast<const double2*>(head + kColBytes)[2 * r + 1];
            } else {
                r01[q].x = reinterpret_cast<const double*>(head + kColBytes)[r];
                r01[q].y = 0.0;
                r23[q] = make_double2(1.0, 1.0);
            }
        }

        // v[q][k]: the pixel (low 3 bytes) of row q, column k of this thread
        unsigned v[kRowsPerThread][4];
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) {
            const Sep1Slot& S = sl[s];
            const bool right = ONE ? right_lens : (s != 0);  // right half of a double image: mirrored columns
            const int w = DBL ? (right ? a.src.wr : a.src.wl) : a.src.W;
            const double cx = DBL ? (right ? a.src.cxr : a.src.cxl) : a.src.cx;
            unsigned g[kRowsPerThread][4];
            if (S.nbox == 0) {  // block-uniform: nothing of this lens is visible from the tile
#pragma unroll
                for (int q = 0; q < kRowsPerThread; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) g[q][k] = 0u;
            } else if (staged) {
                // byte offset inside the stage buffer = py * pitch + px * 3 + (128 + rectangle start - origin)
                const int base = kHead + (s ? sl[0].rect : 0) - S.origin;
                if (S.all_valid) {
#pragma unroll
                    for (int q = 0; q < kRowsPerThread; ++q)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            double fx, fy;
                            camera_fxy(cs[k].x, cs[k].y, s ? r01[q].y : r01[q].x, cy, cx, fx, fy);
                            int px = trunc_abs(fx);
                            if (right) px = a.src.W - 1 - px;
                            const int off = trunc_abs(fy) * S.pitch + (px * 3 + base);
                            g[q][k] = sep1_pick<!DBL || ONE>(stage_sa, off);
                        }
                } else {
#pragma unroll
                    for (int q = 0; q < kRowsPerThread; ++q)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            double fx, fy;
                            camera_fxy(cs[k].x, cs[k].y, s ? r01[q].y : r01[q].x, cy, cx, fx, fy);
                            int px = trunc_abs(fx);
                            if (right) px = a.src.W - 1 - px;
                            int off = trunc_abs(fy) * S.pitch + (px * 3 + base);
                            off = inside_image(fx, fy, w, a.src.H) ? off : 0;  // 0: the buffer's zero bytes
                            g[q][k] = sep1_pick<!DBL || ONE>(stage_sa, off);
                        }
                }
            } else {
                // footprint too large to stage (a pole of the source): straight from global memory
#pragma unroll
                for (int q = 0; q < kRowsPerThread; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        double fx, fy;
                        camera_fxy(cs[k].x, cs[k].y, s ? r01[q].y : r01[q].x, cy, cx, fx, fy);
                        int px = trunc_abs(fx);
                        if (right) px = a.src.W - 1 - px;
                        const int py = trunc_abs(fy);
                        g[q][k] = inside_image(fx, fy, w, a.src.H) ? pick_px_global(frame, py * a.src_pitch + px * 3) : 0u;
                    }
            }
            if (s == 0) {
#pragma unroll
                for (int q = 0; q < kRowsPerThread; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[q][k] = g[q][k];
            } else {
#pragma unroll
                for (int q = 0; q < kRowsPerThread; ++q) {
                    if (r23[q].x == 1.0 && r23[q].y == 1.0) {  // outside the blend band: exact byte add
#pragma unroll
                        for (int k = 0; k < 4; ++k) v[q][k] = __vadd4(v[q][k], g[q][k]);
                    } else {
                        // blend band: the guarded fixed-point blend of pb_tiled.cuh where the row's weights
                        // qualify, the float64 expression otherwise
                        unsigned ia, ib;
                        if (fix_weights(r23[q].x, r23[q].y, ia, ib)) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) v[q][k] = blend_px_fix(v[q][k], ia, r23[q].x, g[q][k], ib, r23[q].y);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) v[q][k] = blend_px_weighted(v[q][k], r23[q].x, g[q][k], r23[q].y);
                        }
                    }
                }
            }
        }

        const unsigned o = out_sa + ob * kOutTileBytes;
#pragma unroll
        for (int q = 0; q < kRowsPerThread; ++q) store_quad_sa(o + q * kRowGroups * kOutRowBytes, v[q]);
        ptx::fence_async_smem();
        if (tid < NSLOT) ring[((it + NB + 1) & 7) * 2 + tid] = pre;
        // The loads of tile it + 2 and the store of this tile are issued by lane 0 of warp (it mod 8):
        // rotating the duty spreads its cost over the warps instead of making one warp late at
        // every barrier.  Bulk async-groups belong to the issuing thread, so the thread that
        // stored the previous tile waits for that store to have left the other output tile.
        if (tid == (((it - 1) & 7) << 5) && it > 0) ptx::bulk_wait_read0();
        __syncthreads();
        if (tid == ((it & 7) << 5)) {
            if (u + NB * G < n_tiles) issue((it + NB) & 7, b);
            ptx::tma_store_3d_hint(&a.dst_map, x0 * 3, y0, 0, out_tiles + ob * kOutTileBytes, drop);
            ptx::bulk_commit();
        }
    }
    if ((tid & 31) == 0) ptx::bulk_wait_read0();  // every issuing thread: its stores have left shared memory
}

inline int sep1_smem_bytes(int cap, bool dbl, int n_buffers) {
    const int head = 128 + kTileW * 16 + kTileH * (dbl ? 32 : 8);
    return 2 * kOutTileBytes + n_buffers * (head + cap) + 16 * (int)sizeof(int4) + 8 * n_buffers + 128;
}

}  // namespace pb
