/*
 * b200sparse.h -- C ABI of libb200sparse.so, the B200 (sm_100a) sparse-convolution library behind
 * the MinkowskiEngine-shaped Python surface in dpcr_agb_b200/MinkowskiEngine.
 *
 * This is the drop-in boundary for the hot path of StefOe/DPCR-AGB's MSENet14 / MSENet50.  The
 * reference has no native code for this path: it calls the third-party MinkowskiEngine library
 * through `import MinkowskiEngine as ME` (torch-points3d/torch_points3d/modules/MinkowskiEngine/
 * SENet.py:3, common.py:6, models/instance/minkowski.py:3).  Every entry point below cites the
 * reference call site whose work it performs ("R:" = /root/reference/torch-points3d/torch_points3d/).
 *
 * Conventions
 *   - every function returns 0 (B2S_OK) or a negative error code; b2s_last_error() gives the text
 *     (thread-local).  No exceptions, no caller-visible allocation: the caller owns every buffer.
 *   - all pointers are device pointers on the CURRENT device unless the name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; the library never touches the default stream and
 *     never synchronises, so it is safe from autograd / DDP worker threads.
 *   - coordinates are int32 [N,4] = (batch, x, y, z)  (R:models/instance/minkowski.py:69).
 *   - a kernel map is a neighbour table nbr int32 [K3, N_out]: nbr[k*N_out + o] = in-row i with
 *     coords_in[i] == coords_out[o] + delta_k, or -1.  k = ix + K*iy + K*K*iz (x fastest).
 *   - features are fp32 row-major [N, C].
 *   - row counts: every function that takes a row count (n, n_out, n_query, m ...) also takes a device pointer
 *     `const int32_t* <name>_dev` right after it.  NULL: the host value is exact.  Non-NULL: the host value is the
 *     CAPACITY of the arrays (pitch of neighbour tables, launch bound) and the kernels read the actual count from
 *     device memory, clamped to [0, capacity].  With device counts no launch shape depends on the data, so a whole
 *     training step (quantise -> maps -> forward -> backward -> optimiser) is captured once into a CUDA graph and
 *     replayed without any host synchronisation (dpcr_agb_b200/graph_step.py).
 */
#ifndef B200SPARSE_H_
#define B200SPARSE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_OK 0
#define B2S_EINVAL (-1)
#define B2S_ENOMEM (-2)
#define B2S_ECUDA (-3)
#define B2S_EOVERFLOW (-4)

#if defined(__GNUC__)
#define B2S_API __attribute__((visibility("default")))
#else
#define B2S_API
#endif

typedef void* b2s_stream_t;

/* ---------------------------------------------------------------- library ------------------- */
B2S_API const char* b2s_last_error(void);
B2S_API int32_t b2s_version(void);
/* 0 iff the current device is compute capability 10.x (B200); B2S_ECUDA otherwise. */
B2S_API int32_t b2s_device_check(void);
/* Diagnostic: overrides one launch-tuning knob of the convolution kernels for this process (results are unaffected,
 * only tile / pipeline shapes).  Keys: "wg_nbp", "wg_lag", "wg_occ2" (weight-gradient stage size cap in 2 KB blocks,
 * producer run-ahead in stages, 1 = two CTAs per SM), "wg_wv" (half-waves of CTAs the row range is split over),
 * "wg_ca" / "tc_ca" (1 = L1-allocating gathers), "tc_occ1" (1 = one CTA per SM), "tc_rot" (1 = every output tile
 * starts its kernel-offset loop at a different offset), "tc_m256" (0 off, 1 = M = 256 tiles for 128-wide output tiles,
 * 2 = wherever they can run, 3 = 64- and 128-wide), "tc_ta" (split-bf16 mode: 0 = A operand always staged in shared
 * memory, 1 = 64-wide output tiles of maps with >= 148 x 256 rows gather it into tensor memory, 2 = 128-wide tiles
 * too, 3 = as 2 for maps of any size), "cr_v4" / "cr_cap" (column reductions: 0 = scalar kernel; CTAs
 * per SM).  A value < 0 restores the default.  B2S_EINVAL for an unknown key.  Not part of the reference-facing
 * surface. */
B2S_API int32_t b2s_set_tuning(const char* key, int32_t value);

/* ---------------------------------------------------------------- (a1) voxel quantisation ----
 * R:core/data_transform/grid_transform.py:112-128 (GridSampling3D._process, mode="last").
 * b2s_quantize_points : q = rint_half_even(fp32(pos) / fp32(size)) as int32 [n,3]   (:116)
 *                       and bounds[0..2] = min(q), bounds[3..5] = max(q) over the batch.
 * b2s_quantize_count  : marks the occupied cells (plot, z, y, x) of the box lo..lo+dims in a bitmap
 *                       held in `workspace` and ranks them; *num_voxels_dev = number of voxels, or
 *                       -1 if a point falls outside the box.  Row order of the result is the order
 *                       of torch.unique(cluster) per plot, plots concatenated (:117-121).
 * b2s_quantize_fill   : representative of a voxel = the point with the largest position in the
 *                       shuffled order (`order[j]` = original index of the j-th shuffled point, NULL =
 *                       identity): consecutive_cluster's last-write-wins (:121).  Writes
 *                       out_coords int32 [M,4] (plot, x, y, z) and out_src int32 [M] (original index).
 */
B2S_API int32_t b2s_quantize_points(const float* pos, int64_t n, const int32_t* n_dev, float size, int32_t* qcoords,
                                    int32_t* bounds, b2s_stream_t stream);
B2S_API int64_t b2s_quantize_workspace_bytes(int64_t num_plots, const int32_t* dims_host);
B2S_API int32_t b2s_quantize_count(const int32_t* qcoords, const int32_t* plot_of_point, int64_t n, const int32_t* n_dev,
                                   int32_t num_plots, const int32_t* lo_host, const int32_t* dims_host, void* workspace,
                                   int64_t workspace_bytes, int32_t* num_voxels_dev, b2s_stream_t stream);
B2S_API int32_t b2s_quantize_fill(const int32_t* qcoords, const int32_t* plot_of_point, const int32_t* order, int64_t n,
                                  const int32_t* n_dev, int32_t num_plots, const int32_t* lo_host,
                                  const int32_t* dims_host, void* workspace, int64_t num_voxels,
                                  const int32_t* num_voxels_dev, int32_t* out_coords, int32_t* out_src,
                                  b2s_stream_t stream);
/* out[r, :] = in[idx[r], :]  -- the per-point tensors gathered at the representative (:64-66). */
B2S_API int32_t b2s_gather_rows(const float* in, const int32_t* idx, int64_t m, const int32_t* m_dev, int32_t c,
                                float* out, b2s_stream_t stream);

/* ---------------------------------------------------------------- (f2) input transforms -------
 * The arithmetic transforms the reference applies per sample on the CPU before the quantiser
 * (R:../conf/data/instance/NFI/transforms/sparse-xy.yaml:105-152), on a collated batch (points of a plot contiguous):
 * b2s_plot_transform : ScalePos(div) + MoveCenterPosPerSample + StartZFromZero
 *                      (R:core/data_transform/transforms.py:590-598, 722-739, 766-769): out = pos / scale + center,
 *                      z -= min z of the plot (minz_scratch: uint32 [num_plots]); keep[i] = 1 iff (x, y) lies inside
 *                      the polygon (Polygon2dExtend :1489-1496, matplotlib's crossings test in double; <= 16 vertices
 *                      as (x, y) pairs, 0 vertices = keep all).
 * b2s_compact_points : stable removal of the points with keep == 0 (apply_mask :1090-1095); index_scratch int32 [n],
 *                      scan_workspace of b2s_scan_workspace_bytes(n); *num_kept_dev = surviving points.
 * b2s_select_by_rank : MaxPoints (:1772-1791, choice = randperm(n)[:num]): rank[i] = position of point i in its plot's
 *                      permutation; rows land in permutation order at offsets[plot] + rank (offsets int32
 *                      [num_plots + 1] written here from the per-plot point counts).
 * b2s_point_features : x = [1, z, ||(x, y) - centre + 1e-6||]  (AddOnes, XYZFeature(z), AddXYDistanceToCenter,
 *                      AddFeatsByKeys: R:core/data_transform/features.py:307-383).
 * b2s_coords_augment : RandomCoordsFlip(ignored z) + ShiftVoxels on quantised coordinates (transforms.py:1046-1054,
 *                      sparse_transforms.py:49-55): aug_dev int32 [num_plots, 5] = (flip_x, flip_y, shift_x, shift_y,
 *                      shift_z) per plot; a flip maps c -> max_c(plot) - c; max_scratch int32 [num_plots, 2].
 */
B2S_API int32_t b2s_plot_transform(const float* pos, const int32_t* plot_of_point, int64_t n, const int32_t* n_dev,
                                   int32_t num_plots, const float* scale_host, const float* center_host,
                                   const double* polygon_host, int32_t num_vertices, uint32_t* minz_scratch,
                                   float* out_pos, int32_t* keep, b2s_stream_t stream);
B2S_API int32_t b2s_compact_points(const float* pos, const int32_t* plot_of_point, const int32_t* keep, int64_t n,
                                   const int32_t* n_dev, int32_t* index_scratch, void* scan_workspace, float* out_pos,
                                   int32_t* out_plot, int32_t* num_kept_dev, b2s_stream_t stream);
B2S_API int32_t b2s_select_by_rank(const float* pos, const int32_t* plot_of_point, const int32_t* rank, int64_t n,
                                   const int32_t* n_dev, int32_t num_plots, int32_t num, const int32_t* counts,
                                   int32_t* offsets, float* out_pos, int32_t* out_plot, b2s_stream_t stream);
B2S_API int32_t b2s_point_features(const float* pos, int64_t n, const int32_t* n_dev, float center_x, float center_y,
                                   float* feats, b2s_stream_t stream);
B2S_API int32_t b2s_coords_augment(int32_t* coords, int64_t m, const int32_t* m_dev, int32_t num_plots,
                                   const int32_t* aug_dev, int32_t* max_scratch, b2s_stream_t stream);

/* ---------------------------------------------------------------- (a2,a3) coordinate maps ----
 * R:models/instance/minkowski.py:74 (ME.SparseTensor -> hash build) and every stride-2 op
 * (R:modules/MinkowskiEngine/SENet.py:53,94-97; resnet_block.py:48-50).
 * The table is an open-addressing hash of `capacity` 16-byte entries {uint64 key, int32 row, pad}.
 * insert : key = pack(batch, floor(c / ts) * ts); value = smallest inserting row (first occurrence).
 *          info_dev (int32[4]): [0] number of unique keys, [1] 1 if a coordinate is out of range,
 *          [2] largest batch id seen (-1 if n == 0), [3] reserved.
 *          rank[i] = out row of in-row i if i is the first occurrence of its key, else -1;
 *          slot[i] = table slot of row i's key (scratch for _fill).
 * fill   : writes out_coords [M,4] in first-occurrence order, rewrites table values to out rows and
 *          (optionally) in2out[i] = out row of in-row i.
 */
B2S_API int64_t b2s_hash_capacity(int64_t n);
B2S_API int64_t b2s_scan_workspace_bytes(int64_t n);
B2S_API int32_t b2s_coordmap_insert(const int32_t* coords, int64_t n, const int32_t* n_dev, const int32_t* ts_host,
                                    void* table, int64_t capacity, int32_t* slot, int32_t* rank, int32_t* info_dev,
                                    void* scan_workspace, b2s_stream_t stream);
/* out_capacity = rows allocated at out_coords; unique keys ranked beyond it are dropped (their table value becomes
 * -1) -- the caller compares info_dev[0] with its capacity afterwards. */
B2S_API int32_t b2s_coordmap_fill(const int32_t* coords, int64_t n, const int32_t* n_dev, const int32_t* ts_host,
                                  void* table, int64_t capacity, const int32_t* slot, const int32_t* rank,
                                  int32_t* out_coords, int64_t out_capacity, int32_t* in2out, b2s_stream_t stream);

/* ---------------------------------------------------------------- (a4) kernel maps -----------
 * Implicit in every MinkowskiConvolution / MinkowskiMaxPooling call (same call sites).
 * nbr[k*n_query + q] = table row of (query[q] + sign * delta_k) or -1;
 * delta = (i - K/2) * step for odd K, i * step for even K; step = dilation * tensor_stride_in.
 * sign=+1: forward map (query = out coords, table = in map).  sign=-1 with query = in coords and the
 * table of the OUT map gives the transposed map (dgrad of strided convs, ConvolutionTranspose).
 * pair_counts / pairs_fill derive MinkowskiEngine's pair-list form (per offset, sorted by out row).
 */
B2S_API int32_t b2s_kernel_map(const int32_t* query_coords, int64_t n_query, const int32_t* n_query_dev,
                               const void* table, int64_t capacity, const int32_t* kernel_size_host,
                               const int32_t* step_host, int32_t sign, int32_t* nbr, b2s_stream_t stream);
/* Same table through the quantiser's occupancy index (the `workspace` of b2s_quantize_count/_fill with the same
 * num_plots / lo / dims) instead of the hash: valid when the rows of the looked-up map ARE the quantiser's output rows
 * (unique, sorted by (plot, z, y, x)); row = prefix popcount rank.  Bit-identical to b2s_kernel_map. */
B2S_API int32_t b2s_kernel_map_dense(const int32_t* query_coords, int64_t n_query, const int32_t* n_query_dev,
                                     const void* quantize_workspace, int32_t num_plots, const int32_t* lo_host,
                                     const int32_t* dims_host, const int32_t* kernel_size_host,
                                     const int32_t* step_host, int32_t sign, int32_t* nbr, b2s_stream_t stream);
/* x-LINE form of a stride-1 kernel map over the quantiser's rows (the k7 stem of R:modules/MinkowskiEngine/SENet.py:49-52
 * reads its 343 offsets as 49 lines of 7 x-consecutive cells): lines[l*n_query + q], l = iy + K[1]*iz, holds
 * (base << 8) | mask -- mask bit ix set <=> offset (ix, iy, iz) of row q has a neighbour, and because the rows are
 * sorted by (plot, z, y, x) those neighbours are the CONSECUTIVE rows base, base+1, ...:
 * nbr[(ix + K[0]*l), q] = base + popc(mask & ((1 << ix) - 1)).  Needs kernel_size[0] <= 8, step[0] == 1 and fewer than
 * 2^24 rows (B2S_EOVERFLOW otherwise).  12 % of the bytes of the [K^3, n] table for K = 7. */
B2S_API int32_t b2s_kernel_map_lines(const int32_t* query_coords, int64_t n_query, const int32_t* n_query_dev,
                                     const void* quantize_workspace, int32_t num_plots, const int32_t* lo_host,
                                     const int32_t* dims_host, const int32_t* kernel_size_host,
                                     const int32_t* step_host, uint32_t* lines, b2s_stream_t stream);
B2S_API int32_t b2s_kernel_map_pair_counts(const int32_t* nbr, int32_t k3, int64_t n_query, int32_t* counts,
                                   b2s_stream_t stream);
B2S_API int32_t b2s_kernel_map_pairs_fill(const int32_t* nbr, int32_t k3, int64_t n_query, const int64_t* offsets,
                                  int32_t* in_idx, int32_t* out_idx, b2s_stream_t stream);

/* ---------------------------------------------------------------- (a6-a8) convolution --------
 * R:modules/MinkowskiEngine/SENet.py:49-52,94-97; resnet_block.py:48-54,95-107; common.py:219-221.
 * b2s_conv_gather_gemm : y[o,:] = bias + sum_k x[nbr[k,o],:] * B_k          (output-stationary)
 *     w_layout bit 0 clear: w is [K3, c_in, c_out]  (forward with the stored kernel)
 *     w_layout bit 0 set  : w is [K3, c_out, c_in]  (dgrad: x = grad_out, nbr = transposed map, w = kernel)
 *     w_layout bit 1 set  : B_k is taken from w[K3-1-k] -- for point-symmetric maps (stride 1, odd K) the
 *                           transposed table is the forward table with k reversed, so dgrad reuses nbr.
 *     w_layout bit 2 set  : x is already TF32-representable (see b2s_round_tf32) -- skip the internal rounding pass.
 *     w_layout bit 4 set  : `workspace` already holds the weight image of this (w, w_layout bits 0-1), built ahead
 *                           with b2s_conv_weight_image (needs bit 2 and c_in > 4): the call launches the convolution
 *                           kernel only.  Weights change once per optimiser step, so a trainer builds the images of
 *                           every layer at the start of the step on a side stream, off the critical path.  For
 *                           shapes that the tensor-memory-operand kernel may take (c_in = 32 * 2^s, c_out a
 *                           multiple of 64 but not of 256) a prebuilt image holds two forms back to back -- the
 *                           plain one and the one in that kernel's channel order -- because the builder does not
 *                           know the row count of the call that will use it; b2s_conv_weight_image_bytes counts both.
 *     k3 == 1 and nbr == NULL means the identity map (the K=1, stride=1 `use_mm` case).
 *     impl: 0 = auto, 1 = SIMT fp32 reference kernel, 2 = tcgen05 kind::tf32 kernel.
 * b2s_conv_wgrad       : gw[k] = sum_o x[nbr[k,o],:]^T gy[o,:]   (gw fp32 [K3, c_in, c_out]);
 *                        flags bit 0: x and gy are already TF32-representable.
 * b2s_colsum           : out[c] = sum_rows x[r,c]                (bias gradient)
 * b2s_round_tf32       : y = x rounded to TF32 (10-bit mantissa, round-to-nearest, ties away: cvt.rna.tf32.f32).
 *                        The tensor-core kernels compute with TF32 operands and fp32 accumulation, and tcgen05 itself
 *                        TRUNCATES fp32 operands; an operand that is consumed by several passes (x by forward and
 *                        wgrad, grad_out by dgrad and wgrad) is rounded once by the caller and flagged pre-rounded,
 *                        otherwise the entry points round internally into `workspace`.
 * b2s_conv_workspace_bytes: prerounded != 0 sizes the workspace for calls that set the pre-rounded flags.
 * col_stats (b2s_conv_gather_gemm, b2s_conv_lines_fwd; nullable): float [b2s_conv_col_stats_elems(n_out, c_out)] that
 *     receives the statistics of the batch norm that follows every convolution of the reference's networks
 *     (ME/SENet.py:49-52, resnet_block.py:48-55), accumulated in the convolution's epilogue while the output tile is in
 *     registers instead of re-reading y: ceil(n_out / 128) partial rows [column sums | column sums of squares], one
 *     per row tile, written with plain stores (thousands of same-address atomics serialise in L2: measured slower than
 *     the pass they replace), then one header float = out rows per partial row.  Launches whose tiles are partial sums
 *     (split-K) and the SIMT path fill the same layout with a reduction over y.  b2s_bn_finalize adds the live partial
 *     rows in fp64 and produces mean / invstd / running statistics exactly as b2s_bn_stats does.
 */
B2S_API int64_t b2s_conv_workspace_bytes(int64_t n_in, int64_t n_out, int32_t c_in, int32_t c_out, int32_t k3,
                                         int32_t prerounded);
B2S_API int64_t b2s_conv_col_stats_elems(int64_t n_out, int32_t c_out);
B2S_API int64_t b2s_conv_weight_image_bytes(int32_t c_in, int32_t c_out, int32_t k3);   /* -1: shape not covered */
B2S_API int32_t b2s_conv_weight_image(const float* w, int32_t c_in, int32_t c_out, int32_t k3, int32_t w_layout,
                                      void* img, int64_t img_bytes, b2s_stream_t stream);
B2S_API int32_t b2s_round_tf32(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* y, b2s_stream_t stream);
B2S_API int32_t b2s_conv_gather_gemm(const float* x, const float* w, const float* bias, const int32_t* nbr, int64_t n_in,
                                     int64_t n_out, const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3,
                                     int32_t w_layout, float* y, void* workspace, int64_t workspace_bytes,
                                     int32_t impl, float* col_stats, b2s_stream_t stream);
B2S_API int32_t b2s_conv_wgrad(const float* x, const float* gy, const int32_t* nbr, int64_t n_in, int64_t n_out,
                               const int32_t* n_out_dev, int32_t c_in, int32_t c_out, int32_t k3, float* gw,
                               void* workspace, int64_t workspace_bytes, int32_t impl, int32_t flags,
                               b2s_stream_t stream);
/* Convolution of a FEW input channels (c_in <= 4: the k7 stem, R:modules/MinkowskiEngine/SENet.py:49-52) through the
 * x-line table of b2s_kernel_map_lines: forward and weight gradient, split-bf16 operand mode only, c_out % 64 == 0.
 * A pipeline stage is one line: the producers read one line word per row, load only the EXISTING neighbours (89 % of
 * the stem's 343 offsets are empty) and store the operand tile to shared memory; no [K^3, n] table is read or built.
 * b2s_conv_lines_supported: 1 when the library is in split-bf16 mode and the shape is covered, else 0 (callers then
 * use b2s_kernel_map_dense + b2s_conv_gather_gemm / b2s_conv_wgrad, which give the same results).
 * x is plain fp32 [n_in, c_in]; gy of b2s_conv_lines_wgrad must be in operand form (b2s_round_tf32 in split-bf16 mode).
 * Same results as b2s_conv_gather_gemm / b2s_conv_wgrad on the expanded table. */
B2S_API int32_t b2s_conv_lines_supported(int32_t c_in, int32_t c_out, const int32_t* kernel_size_host);
B2S_API int64_t b2s_conv_lines_workspace_bytes(int64_t n_in, int32_t c_in, int32_t c_out,
                                               const int32_t* kernel_size_host);
B2S_API int32_t b2s_conv_lines_fwd(const float* x, const float* w, const float* bias, const uint32_t* lines,
                                   int64_t n_in, int64_t n_out, const int32_t* n_out_dev, int32_t c_in, int32_t c_out,
                                   const int32_t* kernel_size_host, float* y, void* workspace, int64_t workspace_bytes,
                                   float* col_stats, b2s_stream_t stream);
B2S_API int32_t b2s_conv_lines_wgrad(const float* x, const float* gy, const uint32_t* lines, int64_t n_in,
                                     int64_t n_out, const int32_t* n_out_dev, int32_t c_in, int32_t c_out,
                                     const int32_t* kernel_size_host, float* gw, void* workspace,
                                     int64_t workspace_bytes, b2s_stream_t stream);
/* dgrad of a STRIDE-2 convolution (gx[i] = sum_k gy[inv[k,i]] W[k]^T) without gathering zero rows.  In the
 * transposed map a fine row has partners only at the offsets whose parity matches its position inside the 2x2x2 coarse
 * cell (27/8 offsets on average).  b2s_parity_plan sorts the fine rows of `coords` by that position into tile-aligned
 * classes: perm int32 [b2s_parity_plan_rows(n)] (out row of every tile row, -1 = padding), bounds int32 [9] (first
 * 128-row tile of every class, [8] = number of tiles), scratch16 = 16 ints.  b2s_conv_dgrad_strided then walks, per
 * tile, only the offsets of its class.  gy must be TF32-representable (b2s_round_tf32); w is the forward kernel
 * [K3, c_x, c_gy]; inv_nbr is the transposed table [K3, n_fine] (b2s_kernel_map with sign = -1).  Same result as
 * b2s_conv_gather_gemm with w_layout = 1 on inv_nbr.  flags bit 0: `workspace` already holds the weight image
 * (b2s_conv_weight_image(w, c_gy, c_x, k3, 1, ...)). */
B2S_API int64_t b2s_parity_plan_rows(int64_t n);
B2S_API int32_t b2s_parity_plan(const int32_t* coords, int64_t n, const int32_t* n_dev, const int32_t* ts_coarse_host,
                                int32_t* perm, int32_t* bounds, int32_t* scratch16, b2s_stream_t stream);
B2S_API int64_t b2s_conv_dgrad_strided_workspace_bytes(int32_t c_gy, int32_t c_x, int32_t k3);
B2S_API int32_t b2s_conv_dgrad_strided(const float* gy, const float* w, const int32_t* inv_nbr, const int32_t* perm,
                                       const int32_t* bounds, int64_t n_fine, const int32_t* n_fine_dev, int32_t c_gy,
                                       int32_t c_x, const int32_t* kernel_size_host, float* gx, void* workspace,
                                       int64_t workspace_bytes, int32_t flags, b2s_stream_t stream);
B2S_API int32_t b2s_colsum(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* out, b2s_stream_t stream);

/* ---------------------------------------------------------------- (a9) max pooling -----------
 * R:modules/MinkowskiEngine/SENet.py:53.  y[o,c] = max_k x[nbr[k,o],c]; arg[o,c] = winning in-row
 * (lowest row on ties, -1 if the out row has no neighbour -> y = 0).  bwd: gx[arg] += gy.
 */
B2S_API int32_t b2s_maxpool_fwd(const float* x, const int32_t* nbr, int64_t n_out, const int32_t* n_out_dev, int32_t c,
                                int32_t k3, float* y, int32_t* arg, float* y_tf32, b2s_stream_t stream);
B2S_API int32_t b2s_maxpool_bwd(const float* gy, const int32_t* arg, int64_t n_in, int64_t n_out,
                                const int32_t* n_out_dev, int32_t c, float* gx, b2s_stream_t stream);

/* ---------------------------------------------------------------- (f3) local sum / average pooling --
 * R:modules/MinkowskiEngine/networks.py:29 (ME.MinkowskiAvgPooling(kernel_size=2, stride=2) of ResNetBase) and
 * MinkowskiSumPooling.  y[o,:] = post_scale[o] * sum_k pre_scale[i] * x[i,:] over i = nbr[k,o] >= 0; both scales
 * nullable.  Average pooling: post_scale = b2s_nbr_inv_counts(nbr) (1 / number of inputs under the kernel, as
 * MinkowskiEngine divides); its backward is the same kernel on the transposed table with pre_scale = that vector.
 */
B2S_API int32_t b2s_sumpool(const float* x, const int32_t* nbr, const float* pre_scale, const float* post_scale,
                            int64_t n_out, const int32_t* n_out_dev, int32_t c, int32_t k3, float* y,
                            b2s_stream_t stream);
B2S_API int32_t b2s_nbr_inv_counts(const int32_t* nbr, int32_t k3, int64_t n, const int32_t* n_dev, float* inv,
                                   b2s_stream_t stream);

/* ---------------------------------------------------------------- (a5,a10,a11) per-plot ops --
 * R:modules/MinkowskiEngine/senet_block.py:43-50; common.py:44-48; SENet.py:63,117.
 * row_batch points at the batch id of row 0 and is read with `row_batch_stride` int32s per row (the
 * coords array itself: stride 4).  scale (nullable) is a per-plot factor, e.g. 1/count for average.
 */
B2S_API int32_t b2s_batch_counts(const int32_t* row_batch, int32_t row_batch_stride, int64_t n, const int32_t* n_dev,
                                 int32_t num_batches, int32_t* counts, b2s_stream_t stream);
B2S_API int32_t b2s_segment_sum(const float* x, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                const int32_t* n_dev, int32_t c, int32_t num_batches, const float* scale, float* y,
                                b2s_stream_t stream);
/* per-plot maximum (MinkowskiGlobalMaxPooling: R:modules/MinkowskiEngine/networks.py:39, PointNet.py:28): rows must be
 * batch-sorted; arg int32 [B, c] = winning row (lowest on ties, -1 for an empty plot -> y = 0); bwd scatters gy to it. */
B2S_API int32_t b2s_segment_max(const float* x, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                const int32_t* n_dev, int32_t c, int32_t num_batches, float* y, int32_t* arg,
                                b2s_stream_t stream);
B2S_API int32_t b2s_segment_max_bwd(const float* gy, const int32_t* arg, int64_t n, int32_t c, int32_t num_batches,
                                    float* gx, b2s_stream_t stream);
B2S_API int32_t b2s_segment_bcast(const float* y, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                  const int32_t* n_dev, int32_t c, const float* scale, float* x_out,
                                  b2s_stream_t stream);
B2S_API int32_t b2s_bcast_mul_fwd(const float* x, const float* y, const int32_t* row_batch, int32_t row_batch_stride,
                                  int64_t n, const int32_t* n_dev, int32_t c, int32_t y_c, float* out,
                                  b2s_stream_t stream);
B2S_API int32_t b2s_bcast_mul_bwd(const float* g, const float* x, const float* y, const int32_t* row_batch,
                                  int32_t row_batch_stride, int64_t n, const int32_t* n_dev, int32_t c,
                                  int32_t num_batches, float* gx, float* gy, b2s_stream_t stream);

/* ---------------------------------------------------------------- (a13,a14,a16) norm / act ---
 * R:modules/MinkowskiEngine/SENet.py:35,51,98; resnet_block.py:51-56,65,73; common.py:41.
 * bn_stats      : batch mean[c] and invstd[c] = 1/sqrt(biased var + eps) over n rows, and (if the
 *                 pointers are non-null) the nn.BatchNorm1d running-statistics update
 *                 running = (1-momentum)*running + momentum*stat with the UNBIASED variance.
 *                 stats_ws = 2*c doubles of scratch.
 * bn_apply      : y = (x-mean)*invstd*gamma+beta, act 0 = none, 1 = exact-erf GELU (fused).
 *                 gamma / beta may be null (affine=False).
 * bn_bwd_reduce : sums[0..c) = sum gy', sums[c..2c) = sum gy'*xhat  (gy' = gy through the fused act);
 *                 these are also grad_beta and grad_gamma.
 * bn_bwd_apply  : training: gx = gamma*invstd*(gy' - sums0/n - xhat*sums1/n); eval: gx = gamma*invstd*gy'.
 * stats_ws      : caller-provided scratch of 2c + 1 doubles (2c accumulators + a ticket with which the last CTA of the
 *                 reduction is elected to finalise: mean / invstd / running buffers, or the fp32 sums).
 * gelu_fwd/bwd  : exact erf GELU on a flat array.
 * TF32 twins     : maxpool_fwd, bn_apply, bn_bwd_apply and add_gelu_fwd take a nullable `*_tf32` output that receives
 *                 the result rounded to TF32 (round-to-nearest), i.e. the operand form b2s_round_tf32 would produce
 *                 for the convolution that consumes it -- the separate rounding pass disappears.  Where the plain
 *                 result has no other consumer (bn_apply, add_gelu_fwd) `y` may be NULL and only the twin is written.
 * gx_colsum      : nullable float [b2s_bn_bwd_colsum_rows(n, c), c] output of bn_bwd_apply: partial rows whose sum over
 *                 the rows is the column sum of gx, i.e. the bias gradient of the convolution in front of the batch
 *                 norm (ME/SENet.py:49-52, resnet_block.py:48-55: conv -> norm), accumulated while gx is written
 *                 instead of re-reading gx in a b2s_colsum launch (one row per block, plain stores, no atomics).
 */
B2S_API int32_t b2s_bn_stats(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float eps, float momentum,
                             float* running_mean, float* running_var, double* stats_ws, float* mean, float* invstd,
                             b2s_stream_t stream);
B2S_API int32_t b2s_bn_finalize(const float* col_stats, int64_t n, const int32_t* n_dev, int32_t c, float eps,
                                float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                                b2s_stream_t stream);
B2S_API int32_t b2s_bn_apply(const float* x, const float* mean, const float* invstd, const float* gamma, const float* beta,
                             int64_t n, const int32_t* n_dev, int32_t c, int32_t act, float* y, float* y_tf32,
                             b2s_stream_t stream);
B2S_API int32_t b2s_bn_bwd_reduce(const float* gy, const float* x, const float* mean, const float* invstd,
                                  const float* gamma, const float* beta, int64_t n, const int32_t* n_dev, int32_t c,
                                  int32_t act, double* stats_ws, float* sums, b2s_stream_t stream);
B2S_API int64_t b2s_bn_bwd_colsum_rows(int64_t n, int32_t c);
/* out[c] = sum over the rows of x [rows, c]: adds the partial rows of gx_colsum (the convolution's bias gradient) */
B2S_API int32_t b2s_sum_rows(const float* x, int64_t rows, int32_t c, float* out, b2s_stream_t stream);
B2S_API int32_t b2s_bn_bwd_apply(const float* gy, const float* x, const float* mean, const float* invstd, const float* gamma,
                                 const float* beta, const float* sums, int64_t n, const int32_t* n_dev, int32_t c,
                                 int32_t act, int32_t training, float* gx, float* gx_tf32, float* gx_colsum,
                                 b2s_stream_t stream);
/* elementwise over n rows of c floats */
B2S_API int32_t b2s_gelu_fwd(const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* y, b2s_stream_t stream);
/* sum = a + b, y = gelu(sum): the residual join out = act(drop_path(out) + residual); backward = gelu_bwd(gy, sum)
 * for both addends (R:modules/MinkowskiEngine/senet_block.py:93-94, resnet_block.py:72-73). */
B2S_API int32_t b2s_add_gelu_fwd(const float* a, const float* b, int64_t n, const int32_t* n_dev, int32_t c, float* sum,
                                 float* y, float* y_tf32, b2s_stream_t stream);
B2S_API int32_t b2s_gelu_bwd(const float* gy, const float* x, int64_t n, const int32_t* n_dev, int32_t c, float* gx,
                             b2s_stream_t stream);

/* ---------------------------------------------------------------- fused SE block tail --------
 * R:modules/MinkowskiEngine/senet_block.py:33-50 (SELayer) + :83-94 (drop path, residual add, activation):
 *     y = gelu( u * (sigmoid(W2 gelu(W1 mean_plot(u) + b1) + b2) * keep[plot]) + res )
 * se_gate_fwd        : the excitation MLP of all plots in one launch; pooled [B,c] (per-plot mean of u, from
 *                      b2s_segment_sum), W1 [h,c], W2 [c,h] (nn.Linear layout), keep [B] = drop-path scale or NULL;
 *                      writes h_pre [B,h], gate [B,c], gate_eff = gate * keep [B,c].
 * gated_add_gelu_fwd : sum = u * gate_eff[plot] + res, y = gelu(sum); y and / or its TF32 twin (see above).
 * gated_add_gelu_bwd : g_res = gy * gelu'(sum), g_u = g_res * gate_eff[plot], g_gate_eff[plot] = sum_rows g_res * u
 *                      (c % 64 == 0).
 * se_gate_bwd        : back through keep, sigmoid, W2, GELU, W1 and the mean: g_pooled [B,c] (already scaled by
 *                      inv_count[plot]), gradients of W1, b1, W2, b2 (written, not accumulated); gz2 [B,c] and
 *                      gh_pre [2,B,h] are caller-provided scratch.
 * bcast_add_         : x[r,:] += y[plot(r),:] in place (adds the pooled branch's gradient to g_u).
 */
B2S_API int32_t b2s_se_gate_fwd(const float* pooled, const float* w1, const float* b1, const float* w2, const float* b2,
                                const float* keep, int32_t num_batches, int32_t c, int32_t h, float* h_pre, float* gate,
                                float* gate_eff, b2s_stream_t stream);
B2S_API int32_t b2s_se_gate_bwd(const float* g_gate_eff, const float* keep, const float* gate, const float* h_pre,
                                const float* pooled, const float* w1, const float* w2, const float* inv_count,
                                int32_t num_batches, int32_t c, int32_t h, float* gz2, float* gh_pre, float* g_pooled,
                                float* gw1, float* gb1, float* gw2, float* gb2, b2s_stream_t stream);
B2S_API int32_t b2s_gated_add_gelu_fwd(const float* u, const float* gate_eff, const float* res, const int32_t* row_batch,
                                       int32_t row_batch_stride, int64_t n, const int32_t* n_dev, int32_t c, float* sum,
                                       float* y, float* y_tf32, b2s_stream_t stream);
B2S_API int32_t b2s_gated_add_gelu_bwd(const float* gy, const float* sum, const float* u, const float* gate_eff,
                                       const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                                       const int32_t* n_dev, int32_t c, int32_t num_batches, float* g_res, float* g_u,
                                       float* g_gate_eff, b2s_stream_t stream);
B2S_API int32_t b2s_bcast_add_(float* x, const float* y, const int32_t* row_batch, int32_t row_batch_stride, int64_t n,
                               const int32_t* n_dev, int32_t c, b2s_stream_t stream);

/* ---------------------------------------------------------------- optimiser ----
 * R:core/optimizer/adabelief.py:90-201 over one flat fp32 parameter buffer, with GradScaler
 * unscale (R:models/base_model.py:241), clip_grad_value_ (:243) and the inf/nan skip fused in.
 * found_inf_dev[0] != 0 -> the step is skipped (GradScaler semantics).  hyper: 16 floats, see csrc/optim.cu; given
 * either on the host (hyper_host, baked into the launch) or on the device (hyper_dev, read by the kernel -- the
 * form a captured CUDA graph needs, since the learning rate changes every step).  Exactly one must be non-NULL.
 */
B2S_API int32_t b2s_grad_check(const float* grad, int64_t numel, float inv_scale, float* found_inf_dev, b2s_stream_t stream);
B2S_API int32_t b2s_adabelief_step(float* param, const float* grad, float* exp_avg, float* exp_avg_var, int64_t numel,
                                   const float* hyper_host, const float* hyper_dev, const float* found_inf_dev,
                                   b2s_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SPARSE_H_ */
