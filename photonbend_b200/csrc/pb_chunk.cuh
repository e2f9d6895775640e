// pb_chunk.cuh -- batches of frames through a separable geometry, source staged as a CHUNK LIST.
//
// remap_tiled_kernel (pb_tiled.cuh) stages the BOUNDING RECTANGLE of a tile's source footprint with
// TMA box loads.  The footprint of a 32 x 64 output tile is an annular sector; where it lies
// diagonally in the source image (the rim of a fisheye circle, the corners of its square) its
// bounding rectangle holds 2-3.3x the bytes the tile reads.  For the tiles of a double-fisheye
// source that see BOTH lenses that was 34 KB per frame for 10 KB of touched pixels, 1.85x their
// algorithmic source bytes out of DRAM, and the shared-memory write port busy with bytes nobody
// reads (profiles/r1_ncu_full_cfg5_16frames.txt).
//
// Here a tile stages exactly the 16-byte chunks its pixels touch (17 KB for such a tile, the
// footprint itself at 16-byte granularity):
//   * once per tile: every thread resolves its 8 pixels per lens (the same float64 table
//     arithmetic as pb_tiled.cuh, bit-identical), marks the chunks they touch in a shared-memory
//     bitmap (one 32-bit word per source row and lens), a block-wide prefix sum packs the marked
//     chunks row-major into a dense list, and every pixel's address inside the packed list is
//     worked out once;
//   * per frame: all 256 threads copy "their" chunks (list entries tid, tid + 256, ...) global ->
//     shared with 16-byte cp.async (LDGSTS), completion on the frame group's mbarrier
//     (cp.async.mbarrier.arrive.noinc); no single thread issues loads for the block, and there is
//     no box geometry to get into uniform registers (what sank the staircase-of-boxes experiment,
//     profiles/experiments/README.md);
//   * gather, blend, shared output tile and the TMA tile store are those of the lean frame loop of
//     pb_tiled.cuh.
// A pixel that straddles two chunks marks both; the second one is the next entry of the same row,
// so "second word = first word + 4" still holds in the packed layout.
#pragma once

#include "pb_tiled.cuh"

namespace pb {

constexpr int kChunkRows = 128;   // source rows per lens a tile's footprint may span
constexpr int kChunkCopies = 6;   // 16-byte copies per thread and frame: up to 1536 chunks = 24 KB per frame

struct alignas(16) ChunkShared {
    uint64_t bar[kMaxGroups];
    unsigned bm[2 * kChunkRows];        // chunks touched, per (lens, source row of the footprint)
    unsigned base[2 * kChunkRows + 1];  // exclusive prefix sum of popc(bm): first list entry of each row
    unsigned warp_sum[8];
};

namespace ptx {
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// the mbarrier gets one arrival from this thread once all its cp.async so far have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar_sa) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar_sa) : "memory");
}
}  // namespace ptx

constexpr int chunk_min_ctas(int src_kind, int cls) {
    return src_kind != PB_KIND_DOUBLE ? 4 : cls == 1 ? 4 : 2;
}

// SRC_KIND: PB_KIND_CAMERA or PB_KIND_DOUBLE; CLS as in remap_tiled_kernel (0 any tile, 1 tiles
// that see one lens at unit weights, 2 the rest).  Output is an un-rotated equirect image
// (separable tables), C = 3.
template <int SRC_KIND, int CLS>
__global__ void __launch_bounds__(kTileThreads, chunk_min_ctas(SRC_KIND, CLS))
remap_chunk_kernel(const __grid_constant__ TiledArgs a) {
    constexpr bool DBL = (SRC_KIND == PB_KIND_DOUBLE);
    constexpr bool ONE = DBL && CLS == 1;
    constexpr int NSLOT = (DBL && !ONE) ? 2 : 1;
    constexpr int S1 = NSLOT - 1;

    extern __shared__ __align__(128) unsigned char smem[];
    // [ out tiles: 2 x 6144 ][ stage area: stage_bytes ][ ChunkShared ]
    unsigned char* out_tiles = smem;
    unsigned char* stages = smem + 2 * kOutTileBytes;
    ChunkShared* sh = reinterpret_cast<ChunkShared*>(stages + a.stage_bytes);

    const int tid = threadIdx.x;
    const int qc = tid & (kQuadsPerRow - 1);
    const int rg = tid >> 3;
    int tile_x, tile_y;
    if (a.tile_list != nullptr) {
        const int t = __ldg(a.tile_list + blockIdx.x);
        tile_y = t / a.tiles_x;
        tile_x = t - tile_y * a.tiles_x;
    } else if (a.raster_band > 0) {
        const int per_band = a.raster_band * a.tiles_x;
        const int band = blockIdx.x / per_band, within = blockIdx.x - band * per_band;
        const int bh = min(a.raster_band, a.tiles_y - band * a.raster_band);
        tile_x = within / bh;
        tile_y = band * a.raster_band + (within - tile_x * bh);
    } else {
        tile_y = blockIdx.x / a.tiles_x;
        tile_x = blockIdx.x - tile_y * a.tiles_x;
    }
    tile_y += a.tile_y0;
    const int x0 = tile_x * kTileW;
    const int y0 = tile_y * kTileH;
    const int ys = y0 - a.tile_y0 * kTileH;

    if (tid == 0) {
        ptx::prefetch_tensormap(&a.dst_map);
#pragma unroll
        for (int g = 0; g < kMaxGroups; ++g) ptx::mbarrier_init(&sh->bar[g], kTileThreads);
        ptx::fence_mbarrier_init();
    }
    sh->bm[tid] = 0u;  // 2 * kChunkRows == kTileThreads entries

    // ------------------------------------------------------------ tables, footprints
    int4 fpv[NSLOT];
    bool right_lens = false;
    if (ONE) {
        const int4 f0 = __ldg(a.tile_fp + (tile_y * a.tiles_x + tile_x) * 2);
        const int4 f1 = __ldg(a.tile_fp + (tile_y * a.tiles_x + tile_x) * 2 + 1);
        right_lens = f0.z == 0;
        fpv[0] = right_lens ? f1 : f0;
    } else {
#pragma unroll
        for (int s = 0; s < NSLOT; ++s) fpv[s] = __ldg(a.tile_fp + (tile_y * a.tiles_x + tile_x) * NSLOT + s);
    }
    double2 cs[4], r01[kRowsPerThread];
    double wrow[kRowsPerThread][2];
    {
        const double2* __restrict__ col = reinterpret_cast<const double2*>(a.sep1_col) + tile_x * kTileW;
#pragma unroll
        for (int k = 0; k < 4; ++k) cs[k] = __ldg(col + k * 8 + qc);
#pragma unroll
        for (int q = 0; q < kRowsPerThread; ++q) {
            const int r = tile_y * kTileH + rg + q * kRowGroups;
            wrow[q][0] = wrow[q][1] = 1.0;
            if (ONE) {
                r01[q].x = __ldg(a.sep1_row + 4 * r + (right_lens ? 1 : 0));
                r01[q].y = 0.0;
            } else if (DBL) {
                const double2* __restrict__ row = reinterpret_cast<const double2*>(a.sep1_row) + 2 * r;
                r01[q] = __ldg(row);
                const double2 r23 = __ldg(row + 1);
                wrow[q][0] = r23.x;
                wrow[q][1] = r23.y;
            } else {
                r01[q].x = __ldg(a.sep1_row + r);
                r01[q].y = 0.0;
            }
        }
    }
    bool unit_weights = true;
    if (DBL && !ONE) {
        bool mine = true;
#pragma unroll
        for (int q = 0; q < kRowsPerThread; ++q) mine = mine && wrow[q][0] == 1.0 && wrow[q][1] == 1.0;
        unit_weights = __syncthreads_and(mine);  // (also: barrier init and the cleared bitmap are visible)
    } else {
        __syncthreads();
    }

    // ------------------------------------------------------------ resolve: (row, byte) of every pixel, chunk bitmap
    // key = row within the footprint << 16 | byte within the footprint's rows; -1: no source pixel
    int key[NSLOT][kPxPerThread];
    bool fits = true;  // block-uniform: the footprints fit the bitmap (else: gathers from global memory)
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
        const bool right = ONE ? right_lens : (s != 0);
        const int w = DBL ? (right ? a.src.wr : a.src.wl) : a.src.W;
        const double cx = DBL ? (right ? a.src.cxr : a.src.cxl) : a.src.cx;
        const int by0 = fpv[s].x, xb0 = fpv[s].y, nbox = fpv[s].z;
        const bool all_valid = fpv[s].w & 1;
        if (nbox * kBoxRows > kChunkRows || stage_units(fpv[s].w >> 1) > kMaxStageUnits) fits = false;
        if (nbox == 0 || !fits) {
#pragma unroll
            for (int p = 0; p < kPxPerThread; ++p) key[s][p] = -1;
            continue;
        }
        unsigned* bm = sh->bm + s * kChunkRows;
#pragma unroll
        for (int q = 0; q < kRowsPerThread; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                double fx, fy;
                camera_fxy(cs[k].x, cs[k].y, s ? r01[q].y : r01[q].x, a.src.cy, cx, fx, fy);
                int px = trunc_abs(fx);
                if (right) px = a.src.W - 1 - px;
                const int r = trunc_abs(fy) - by0, xb = px * 3 - xb0;
                const bool ok = all_valid || inside_image(fx, fy, w, a.src.H);
                key[s][q * 4 + k] = ok ? (r << 16) | xb : -1;
                if (ok) atomicOr(bm + r, (1u << (xb >> 4)) | (1u << ((xb + 2) >> 4)));
            }
    }
    __syncthreads();

    // ------------------------------------------------------------ pack: prefix sum over the rows' chunk counts
    const unsigned my_bits = sh->bm[tid];
    const unsigned my_cnt = __popc(my_bits);
    unsigned incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) sh->warp_sum[tid >> 5] = incl;
    __syncthreads();
    unsigned before = 0, n_chunks = 0;
#pragma unroll
    for (int wdx = 0; wdx < 8; ++wdx) {
        const unsigned v = sh->warp_sum[wdx];
        if (wdx < (tid >> 5)) before += v;
        n_chunks += v;
    }
    const unsigned my_base = before + incl - my_cnt;
    sh->base[tid] = my_base;

    const int group_bytes = (128 + (int)n_chunks * 16 + 127) & ~127;
    const int n_groups = min(min(kMaxGroups, a.n_frames), a.stage_bytes / group_bytes);
    if (!fits || n_chunks > (unsigned)(kChunkCopies * kTileThreads) || n_groups < 1) {  // block-uniform
        direct_tile<PB_KIND_EQUIRECT, SRC_KIND, 1>(a, nullptr, nullptr, out_tiles, x0, y0);
        return;
    }
    // the list itself (byte offset of every chunk within a source frame), written row by row into the
    // still unused stage area by the thread that owns the row
    {
        unsigned* list = reinterpret_cast<unsigned*>(stages);
        const int s = tid >> 7, r = tid & (kChunkRows - 1);  // tid = lens * kChunkRows + row
        if (s < NSLOT && my_cnt) {
            const unsigned row_off = (unsigned)(fpv[s].x + r) * (unsigned)a.src_pitch + (unsigned)fpv[s].y;
            unsigned m = my_bits, i = my_base;
            while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                list[i++] = row_off + 16u * c;
            }
        }
    }
    __syncthreads();
    unsigned goff[kChunkCopies];
#pragma unroll
    for (int k = 0; k < kChunkCopies; ++k) {
        const unsigned j = tid + k * kTileThreads;
        goff[k] = j < n_chunks ? reinterpret_cast<const unsigned*>(stages)[j] : 0xffffffffu;
    }

    // every pixel's address inside a frame group: [128 zero bytes][packed chunks]
    const unsigned stages_sa = ptx::smem_addr(stages);
    unsigned adr[NSLOT][kPxPerThread], shf[NSLOT][kPxPerThread];
#pragma unroll
    for (int s = 0; s < NSLOT; ++s)
#pragma unroll
        for (int p = 0; p < kPxPerThread; ++p) {
            const int kx = key[s][p];
            unsigned rel = 0;  // the group's zero bytes
            if (kx >= 0) {
                const int e = s * kChunkRows + (kx >> 16), xb = kx & 0xffff;
                const unsigned idx = sh->base[e] + __popc(sh->bm[e] & ((1u << (xb >> 4)) - 1u));
                rel = 128u + idx * 16u + (unsigned)(xb & 15);
            }
            adr[s][p] = stages_sa + (rel & ~3u);
            shf[s][p] = rel << 3;
        }
    unsigned wfix[kRowsPerThread][2];
    int wmode[kRowsPerThread];
#pragma unroll
    for (int q = 0; q < kRowsPerThread; ++q) {
        wfix[q][0] = wfix[q][1] = 0;
        wmode[q] = 0;
        if (DBL && !ONE && !unit_weights && !(wrow[q][0] == 1.0 && wrow[q][1] == 1.0))
            wmode[q] = fix_weights(wrow[q][0], wrow[q][1], wfix[q][0], wfix[q][1]) ? 1 : 2;
    }
    __syncthreads();  // everybody has read the list: the stage area becomes the ring of frame groups
    if (tid < n_groups * 8) reinterpret_cast<int4*>(stages + (tid >> 3) * group_bytes)[tid & 7] = make_int4(0, 0, 0, 0);

    const unsigned bar_sa = ptx::smem_addr(&sh->bar[0]);
    auto issue_group = [&](int f, int g) {  // every thread: its chunks of frame f into group g
        const unsigned dst = stages_sa + (unsigned)(g * group_bytes) + 128u + (unsigned)tid * 16u;
        const unsigned char* __restrict__ frame = a.src_px + (long long)f * a.src_frame_stride;
#pragma unroll
        for (int k = 0; k < kChunkCopies; ++k)
            if (goff[k] != 0xffffffffu) ptx::cp_async16(dst + k * (kTileThreads * 16), frame + goff[k]);
        ptx::cp_async_arrive_noinc(bar_sa + 8 * g);
    };
    for (int f = 0; f < n_groups; ++f) issue_group(f, f);
    __syncthreads();  // the groups' zero bytes are in place

    const uint64_t drop = ptx::policy_evict_first();
    const unsigned out_sa = ptx::smem_addr(out_tiles) + rg * kOutRowBytes + qc * 12;
    const int n_act = (fpv[0].z > 0) + ((NSLOT == 2 && fpv[S1].z > 0) ? 1 : 0);
    const int first = (fpv[0].z > 0) ? 0 : S1;
    auto frame_loop = [&](auto ACT, auto WGT) {
        constexpr int act = decltype(ACT)::value;   // 1: slot 0 only, 2: slot 1 only, 3: both
        constexpr bool wgt = decltype(WGT)::value;  // some row of the tile has a weighted blend
        int g = 0;
        unsigned parity = 0;
        for (int f = 0; f < a.n_frames; ++f) {
            const unsigned goffs = (unsigned)(g * group_bytes);
            ptx::mbarrier_wait_sa(bar_sa + 8 * g, parity);
            unsigned v[kPxPerThread];
            if (wgt) {
                unsigned w[kPxPerThread];
#pragma unroll
                for (int p = 0; p < kPxPerThread; ++p) {
                    v[p] = (act & 1) ? PB_LEAN_PICK(adr[0][p], goffs, shf[0][p]) : 0u;
                    w[p] = (act & 2) ? PB_LEAN_PICK(adr[S1][p], goffs, shf[S1][p]) : 0u;
                }
#pragma unroll
                for (int q = 0; q < kRowsPerThread; ++q) {
                    if (wmode[q] == 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) v[q * 4 + k] = __vadd4(v[q * 4 + k], w[q * 4 + k]);
                    } else if (wmode[q] == 1) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            v[q * 4 + k] = blend_px_fix(v[q * 4 + k], wfix[q][0], wrow[q][0], w[q * 4 + k], wfix[q][1], wrow[q][1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            v[q * 4 + k] = blend_px_weighted(v[q * 4 + k], wrow[q][0], w[q * 4 + k], wrow[q][1]);
                    }
                }
            } else {
                if (act & 1) {
#pragma unroll
                    for (int p = 0; p < kPxPerThread; ++p) v[p] = PB_LEAN_PICK(adr[0][p], goffs, shf[0][p]);
                }
                if (act & 2) {
#pragma unroll
                    for (int p = 0; p < kPxPerThread; ++p) {
                        const unsigned w = PB_LEAN_PICK(adr[S1][p], goffs, shf[S1][p]);
                        v[p] = (act & 1) ? __vadd4(v[p], w) : w;
                    }
                }
            }
            const unsigned o = out_sa + (f & 1) * kOutTileBytes;
#pragma unroll
            for (int q = 0; q < kRowsPerThread; ++q) store_quad_sa(o + q * kRowGroups * kOutRowBytes, v + q * 4);
            ptx::fence_async_smem();
            if (tid == 0) ptx::bulk_wait_read0();  // stores up to frame f - 1 have left their tiles
            __syncthreads();
            if (f + n_groups < a.n_frames) issue_group(f + n_groups, g);
            if (tid == 0) {
                ptx::tma_store_3d_hint(&a.dst_map, x0 * 3, ys, f, out_tiles + (f & 1) * kOutTileBytes, drop);
                ptx::bulk_commit();
            }
            if (++g == n_groups) {
                g = 0;
                parity ^= 1u;
            }
        }
    };
    if (n_act == 0) {
        // nothing visible (cannot happen for a tile with n_chunks > 0 ... but a tile of black pixels has none)
        frame_loop(std::integral_constant<int, 1>{}, std::false_type{});
    } else if (DBL && !ONE && !unit_weights) {
        if (n_act == 2) frame_loop(std::integral_constant<int, 3>{}, std::true_type{});
        else if (first == 0) frame_loop(std::integral_constant<int, 1>{}, std::true_type{});
        else frame_loop(std::integral_constant<int, 2>{}, std::true_type{});
    } else if (n_act == 2) frame_loop(std::integral_constant<int, 3>{}, std::false_type{});
    else if (first == 0) frame_loop(std::integral_constant<int, 1>{}, std::false_type{});
    else frame_loop(std::integral_constant<int, 2>{}, std::false_type{});
    if (tid == 0) ptx::bulk_wait_read0();
}

inline int chunk_smem_bytes(int stage_bytes) {
    return 2 * kOutTileBytes + stage_bytes + (int)sizeof(ChunkShared) + 128;
}

}  // namespace pb
