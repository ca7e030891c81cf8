/* b200scn -- C-ABI of the B200-native sparse U-ResNet hot path (libb200scn.so).
 *
 * This is the drop-in boundary: the set of entry points a replacement for SparseConvNet's
 * native extension (`sparseconvnet.SCN`, reached from
 * /root/reference/uresnet/models/uresnet_sparse.py:9,19-24 through `import sparseconvnet as scn`)
 * has to provide for the modules that model uses.  Each group below names the scn module /
 * reference call site it serves.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer in the caller's current CUDA
 *     context unless its name ends in `_host`; the last argument is a cudaStream_t passed as
 *     void*.  The library never allocates or frees caller memory and never synchronises.
 *   - return value: 0 = ok, <0 = error (message via b200scn_last_error(), thread-local).
 *   - row counts come in two forms: `n_cap` (host int, sizes the grid and bounds every
 *     access) and `n_dev` (device int32*, may be NULL).  If n_dev != NULL the kernels use
 *     min(*n_dev, n_cap) as the live row count.  That is what lets a whole training step be
 *     captured in one CUDA graph although the number of active sites changes per event.
 *   - feature matrices are row-major fp32 with an explicit row stride (`ld*`, in elements),
 *     so a layer can read or write a column slice of a wider matrix (skip-concat buffers).
 *   - index tables are int32, offset-major: table[k * ld + row], -1 = no entry.
 *   - weights: W[k][ci][co] addressed as W[k*w_sk + ci*w_si + co*w_so]; `kflip` != 0 makes
 *     the kernel use W[K-1-k] for table offset k (dgrad of a submanifold conv).
 *   - site keys: 64-bit, sample << 48 | c0 << 32 | c1 << 16 | c2 (3D) or sample << 48 | c0 << 16 | c1
 *     (2D); coordinates and sample ids must be in [0, 65535).
 */
#ifndef B200SCN_H
#define B200SCN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SCN_OK 0
#define B200SCN_ERR_ARG (-1)
#define B200SCN_ERR_CUDA (-2)
#define B200SCN_ERR_UNSUPPORTED (-3)

/* coordinate dtypes accepted by b200scn_input_build */
#define B200SCN_I32 0
#define B200SCN_I64 1
#define B200SCN_F32 2
#define B200SCN_F64 3

/* math modes of the gather-GEMM kernels */
#define B200SCN_MATH_FP32 0   /* FFMA, exact fp32 products                                  */
#define B200SCN_MATH_TF32X3 1 /* mma.sync tf32, 3-term split: fp32-equivalent accuracy       */
#define B200SCN_MATH_TF32 2   /* mma.sync tf32, single pass: ~1e-3 relative                  */

const char* b200scn_last_error(void);
int b200scn_version(void);
/* kernels launched by this library in this process so far (bench.py's gpu_launches evidence) */
int64_t b200scn_launch_count(void);
/* number of SMs of the current device (grid sizing); <0 on error */
int b200scn_sm_count(void);

/* ---------------------------------------------------------------------------------------
 * (i) metadata / rulebook  -- replaces scn's CPU Metadata: InputLayer rules
 * (uresnet_sparse.py:20), SubmanifoldConvolution rules (:21), Convolution/Deconvolution
 * rules inside scn.UNet (:22).
 * ------------------------------------------------------------------------------------- */

/* hash slots (power of two, >= 2 * n_cap) the caller must allocate: keys uint64[slots], vals int32[slots] */
int64_t b200scn_hash_slots(int64_t n_cap);
/* bytes of scratch the rulebook calls need for n_cap rows */
int64_t b200scn_rulebook_scratch_bytes(int64_t n_cap);

/* InputLayer(mode=3) rules.  coords [n_cap, D+1] (or [n_cap, D] when has_sample_col == 0) of
 * dtype coord_dtype, row stride ld_coords elements.
 * Outputs: in2row[n_cap]            row of every input point (first-occurrence order)
 *          ukeys[n_cap]             key of every active row
 *          n_act_dev[1]             number of active rows
 *          first_idx[n_cap]         input index of each active row's first occurrence
 *          seg_off[n_cap + 1], seg[n_cap]  CSR of the *additional* (duplicate) inputs of each
 *                                   row, ascending input index (for in-order mode-3 sums)
 *          err_dev[1]               set non-zero when a coordinate is out of key range
 * The hash table (tkeys, tvals) is left mapping key -> row for later probes. */
int b200scn_input_build(const void* coords, int coord_dtype, int64_t ld_coords, int has_sample_col, int D,
                        int64_t n_cap, const int32_t* n_dev, uint64_t* tkeys, int32_t* tvals, int64_t slots,
                        int32_t* in2row, uint64_t* ukeys, int32_t* n_act_dev, int32_t* first_idx,
                        int32_t* seg_off, int32_t* seg, int32_t* err_dev, void* scratch, int64_t scratch_bytes,
                        void* stream);

/* Convolution(filter == stride == s) rules: next level's active set in discovery order.
 * Outputs: parent[n_cap], kofs[n_cap] (offset code of each fine row inside its parent),
 *          ukeys_next[n_cap], n_next_dev[1], child[K * ld_child] (K = s^D; row of the child
 *          at offset k of each coarse row, -1 = none).  (tkeys, tvals) = table of the new level. */
int b200scn_level_down(const uint64_t* ukeys, int64_t n_cap, const int32_t* n_dev, int D, int s, uint64_t* tkeys,
                       int32_t* tvals, int64_t slots, int32_t* parent, uint8_t* kofs, uint64_t* ukeys_next,
                       int32_t* n_next_dev, int32_t* child, int64_t ld_child, void* scratch,
                       int64_t scratch_bytes, void* stream);

/* SubmanifoldConvolution(filter k, odd) rules as a neighbour table nbr[K * ld], K = k^D,
 * offset index enumerating the box with the last dimension fastest. */
int b200scn_subm_table(const uint64_t* ukeys, int64_t n_cap, const int32_t* n_dev, int D, int k,
                       const uint64_t* tkeys, const int32_t* tvals, int64_t slots, int32_t* nbr, int64_t ld,
                       void* stream);

/* Grouped pair lists (GPL): the rulebook of one (level, layer type) in the form the tensor-core
 * convolution consumes -- per tile of `tile_rows` (32, 64 or 128) output rows, the rule pairs of every
 * offset packed into groups of 16 (csrc/conv_gpl.cu).  Built once per level, shared by every forward /
 * dgrad use.  Source of the pairs: `table` (offset-major, as above), or parent/kofs (Deconvolution-style:
 * pair (parent[p] -> p) at offset kofs[p]), or neither (identity, K == 1).
 * Record format: g_rec[group][8] int4; entry j = { in row of pair j, in row of pair j+8,
 * out(j) | out(j+8) << 8 | k << 16, 0 } with out = row inside the tile (tile_rows = padding slot).
 * stage 0 -> tile_off[tiles + 1] (scanned group offsets), total_dev[1];
 * stage 1 -> g_rec[128 B * cap_groups] at those offsets (tile t = groups tile_off[t] .. tile_off[t+1]);
 * stage 2 -> single pass, fixed-stride slots: tile_off[tiles] = begin, tile_end[tiles] = end of every tile,
 *            cap_groups >= b200scn_gpl_max_groups(); no scan and no read-back (CUDA-graph regime).
 * koff (optional, stages 1 and 2): int32[tiles * (K + 1)], first group of every offset inside each tile
 * (entry K = end of the tile) -- the index b200scn_wgrad_gpl walks.
 * layout: 0 = every offset's pairs compacted into groups, offsets ascending (what every kernel reads).
 *         1 = centre first, submanifold tables only, no koff: the centre offset comes first as one group per 16-row
 *         sub-tile whose slot j IS output row 16 s + j; byte 3 of word 2 carries flag 1 ("C = 0"), telling the
 *         large-input kernel to start that sub-tile's sums in registers instead of reading its accumulator.  The
 *         other offsets follow as in layout 0.  Only the large-input kernel of b200scn_conv_gpl reads these
 *         lists: ask b200scn_conv_gpl_layout. */
int b200scn_gpl_tile_rows(int64_t n); /* recommended tile_rows for a level of n rows */
int64_t b200scn_gpl_max_groups(int64_t n_cap, int K, int tile_rows);
int64_t b200scn_gpl_scratch_bytes(int64_t n_cap);
int b200scn_gpl_build(int stage, const int32_t* table, int64_t ld, const int32_t* parent, const uint8_t* kofs, int K,
                      int64_t n_cap, const int32_t* n_dev, int tile_rows, int layout, int32_t* tile_off,
                      int32_t* tile_end, int32_t* total_dev, void* g_rec, int64_t cap_groups, int32_t* koff,
                      void* scratch, int64_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (ii) gather-GEMM(-scatter) convolutions -- SubmanifoldConvolution / Convolution /
 * Deconvolution / NetworkInNetwork forward, dgrad, wgrad.
 * ------------------------------------------------------------------------------------- */

/* out[o, :] (+)= sum_k in[table[k][o], :] @ W[k]     (table == NULL: K must be 1, identity)
 * serves: subm fwd; subm dgrad (kflip=1, transposed weight strides); Convolution fwd (child
 * table); Deconvolution dgrad (child table, transposed); NetworkInNetwork fwd/dgrad. */
int b200scn_conv_table(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                       int kflip, const int32_t* table, int64_t ld, int K, int64_t n_out_cap,
                       const int32_t* n_out_dev, int Cin, int Cout, float* out, int64_t ldo, int accumulate,
                       int math_mode, void* stream);

/* the same contraction driven by a grouped pair list (tensor-core math modes, channels % 16 == 0);
 * tile_end == NULL: scanned lists (stage 0 + 1), else fixed-stride lists (stage 2);
 * n_out_expect: expected live row count when n_out_cap is a capacity (launch heuristics only; 0 = unknown);
 * layout: the layout the list was built with */
/* Wp: optional fragment-ordered copy of W for this direction (b200scn_pack_weights with the same strides
 * and kflip), or NULL.  res (optional, row stride ldr): added to the result -- the shortcut of a residual block. */
int b200scn_conv_gpl(const float* in, int64_t ldi, const float* W, const float* Wp, int64_t w_sk, int64_t w_si,
                     int64_t w_so, int kflip, int tile_rows, int layout, const int32_t* tile_off,
                     const int32_t* tile_end, const void* g_rec, int K, int64_t n_out_cap, const int32_t* n_out_dev, int64_t n_out_expect, int Cin,
                     int Cout, float* out, int64_t ldo, int accumulate, const float* res, int64_t ldr, int math_mode,
                     void* stream);

/* layout (b200scn_gpl_build) the convolution of these shapes wants its pair list in: 0, or 1 (centre first)
 * when the large-input kernel will run (scanned != 0: lists with scanned tile offsets, i.e. tile_end == NULL;
 * submanifold != 0: the list comes from a submanifold neighbour table). */
int b200scn_conv_gpl_layout(int tile_rows, int scanned, int64_t n_out_cap, int Cin, int Cout, int K, int submanifold,
                            int math_mode);

/* Re-order weights into the mma fragment order b200scn_conv_gpl reads with coalesced 16-byte loads.
 * jobs_dev: device array of n_jobs records of int64[10] = {W ptr, out ptr, K, Cin, Cout, w_sk, w_si, w_so, kflip, 0}
 * (Cin/Cout/strides of the DIRECTION the copy serves: forward = (nIn, nOut, nOut, 1), dgrad = (nOut, nIn, 1, nOut));
 * out holds K*Cin*Cout floats; max_elems = largest K*Cin*Cout among the jobs.  One launch for all jobs. */
int b200scn_pack_weights(const int64_t* jobs_dev, int n_jobs, int64_t max_elems, void* stream);

/* out[p, :] = in[parent[p], :] @ W[kofs[p]]
 * serves: Deconvolution fwd; Convolution dgrad (transposed weight strides). */
int b200scn_conv_parent(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                        const int32_t* parent, const uint8_t* kofs, int K, int64_t n_out_cap,
                        const int32_t* n_out_dev, int Cin, int Cout, float* out, int64_t ldo, int math_mode,
                        void* stream);

/* dW[k] (+)= sum_r A[ra(r,k), :]^T @ B[rb(r,k), :]   with, for gather_b == 0,
 *   ra = table[k][r], rb = r  (subm wgrad: A = in, B = dout;  Convolution wgrad: child table)
 * and for gather_b == 1,  ra = r, rb = table[k][r]       (Deconvolution wgrad: child table).
 * dW is dense [K][Ca][Cb].  Deterministic: partial sums per row-chunk land in scratch and
 * are reduced in fixed order.  table == NULL: K == 1, identity (NetworkInNetwork wgrad). */
int64_t b200scn_wgrad_scratch_bytes(int64_t n_cap, int K, int Ca, int Cb);
int b200scn_wgrad_table(const float* A, int64_t lda, const float* B, int64_t ldb, const int32_t* table,
                        int64_t ld, int K, int gather_b, int64_t n_cap, const int32_t* n_dev, int Ca, int Cb,
                        float* dW, int accumulate, int math_mode, void* scratch, int64_t scratch_bytes,
                        void* stream);

/* Weight gradients over a grouped pair list (tensor-core math modes, channels % 16 == 0): the list must be the
 * one the layer's FORWARD uses (its in rows index A = layer input, its out rows index B = output gradient).
 * subm != 0: offset K/2 is the dense centre of a submanifold convolution (gets extra CTAs).  Deterministic. */
int64_t b200scn_wgrad_gpl_scratch_bytes(int K, int Ca, int Cb);
int b200scn_wgrad_gpl(const float* A, int64_t lda, const float* B, int64_t ldb, int tile_rows, const void* g_rec,
                      const int32_t* koff, int K, int subm, int64_t n_cap, const int32_t* n_dev, int64_t n_expect,
                      int Ca, int Cb, float* dW, int accumulate, int math_mode, void* scratch, int64_t scratch_bytes,
                      void* stream);

/* tcgen05 / TMEM variant of b200scn_conv_table for channel counts that are multiples of 16 (csrc/sparse_tc.cu):
 * 128 output rows per CTA, contraction over (offset, input channel) in 32-wide blocks as tcgen05.mma kind::tf32,
 * accumulator in TMEM, a missing table entry = a zero operand row.  Weights come PRE-PACKED (b200scn_pack_weights_tc,
 * one launch for any number of layers: job = int64[12] {W, Wt, Wt_lo, K, Cin, Cout, w_sk, w_si, w_so, kflip, 0, 0};
 * Wt / Wt_lo = b200scn_conv_tc_packed_elems floats each: tf32 high parts and exact remainders, K-major); the
 * offset flip and the transpose of a dgrad live in the packing strides.  res (optional, row stride ldr) is added to
 * the result (residual connections of scn.UNet fused into the convolution).  math_mode TF32X3 or TF32.
 * n_expect: rows the level typically holds when n_cap is only a buffer capacity (0 = unknown; launch heuristics). */
int64_t b200scn_conv_tc_packed_elems(int K, int Cin, int Cout);
int b200scn_conv_tc_supported(int K, int Cin, int Cout);
int b200scn_pack_weights_tc(const int64_t* jobs_dev, int n_jobs, int64_t max_elems, void* stream);
int b200scn_conv_tc(const float* in, int64_t ldi, const float* Wt, const float* Wt_lo, const int32_t* table,
                    int64_t ld_t, int K, int64_t n_cap, const int32_t* n_dev, int64_t n_expect, int Cin, int Cout,
                    float* out, int64_t ldo, const float* res, int64_t ldr, int math_mode, void* stream);

/* tcgen05 / TMEM weight gradient for the same layer shapes: dW[k][ci][co] (+)= sum_o in[table[k][o]][ci] dout[o][co]
 * (gather on the input side: SubmanifoldConvolution, Convolution, NetworkInNetwork).  Per offset a GEMM that
 * contracts over the rows, so both operands are MN-major tiles; the grid splits the flattened (offset, channel)
 * axis into groups of 128 and the rows into splits whose partial blocks (scratch) are folded in fixed order.
 * dW is contiguous [K][Cin][Cout].  n_expect as for b200scn_conv_tc; pass the same value to the scratch query. */
int64_t b200scn_wgrad_tc_scratch_bytes(int64_t n_rows, int K, int Cin, int Cout);
int b200scn_wgrad_tc(const float* in, int64_t ldi, const float* dout, int64_t ldd, const int32_t* table, int64_t ld_t,
                     int K, int64_t n_cap, const int32_t* n_dev, int64_t n_expect, int Cin, int Cout, float* dW,
                     int accumulate, int math_mode, void* scratch, int64_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (iii) fused BatchNorm + (leaky) ReLU -- scn.BatchNormReLU (uresnet_sparse.py:23) and
 * scn.BatchNormLeakyReLU inside scn.UNet.  eps inside the sqrt, `retention` is scn's
 * "momentum" (0.9 keeps 90% of the running value), running_var uses the unbiased variance.
 * ------------------------------------------------------------------------------------- */
int64_t b200scn_bn_scratch_bytes(int64_t n_cap, int C);
int b200scn_bn_fwd_train(const float* x, int64_t ldx, int64_t n_cap, const int32_t* n_dev, int C,
                         const float* gamma, const float* beta, float* running_mean, float* running_var,
                         float eps, float retention, float leak, float* y, int64_t ldy, float* save_mean,
                         float* save_invstd, void* scratch, int64_t scratch_bytes, void* stream);
int b200scn_bn_fwd_eval(const float* x, int64_t ldx, int64_t n_cap, const int32_t* n_dev, int C,
                        const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float leak, float* y, int64_t ldy, void* stream);
/* backward of the fused op; the activation mask is recomputed from x and the saved stats.
 * `train` is a bit set: bit 0 = batch-stat backward (0: affine backward with fixed (mean, invstd));
 * bit 1 = dgamma / dbeta are ADDED to what the buffers hold (a persistent flat gradient buffer), else written. */
int b200scn_bn_bwd(const float* x, int64_t ldx, const float* dy, int64_t lddy, int64_t n_cap,
                   const int32_t* n_dev, int C, const float* gamma, const float* beta, const float* mean,
                   const float* invstd, float leak, int train, float* dx, int64_t lddx, float* dgamma,
                   float* dbeta, void* scratch, int64_t scratch_bytes, void* stream);

/* Single-launch variants for inputs small enough to sit in the SMs' registers (every level of a bs=1 512^3
 * event): statistics, a grid-wide barrier and the apply phase in ONE co-resident grid (<= one CTA per SM), so
 * the matrix is read once and the second launch disappears.  `sync` = two uint32 of caller-owned device memory,
 * zero before the first use (word 0: wrapping arrival counter, back at zero after every call; word 1 reserved), not shared by
 * calls that may run concurrently; NULL, or an input too large (judged from n_expect, the rows a typical event
 * holds at this level, 0 = unknown), falls through to the two-kernel functions above.  Same results contract. */
int b200scn_bn_fwd_train_sync(const float* x, int64_t ldx, int64_t n_cap, const int32_t* n_dev, int64_t n_expect,
                              int C, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, float eps, float retention, float leak, float* y, int64_t ldy,
                              float* save_mean, float* save_invstd, uint32_t* sync, void* scratch,
                              int64_t scratch_bytes, void* stream);
int b200scn_bn_bwd_sync(const float* x, int64_t ldx, const float* dy, int64_t lddy, int64_t n_cap,
                        const int32_t* n_dev, int64_t n_expect, int C, const float* gamma, const float* beta,
                        const float* mean, const float* invstd, float leak, int train, float* dx, int64_t lddx,
                        float* dgamma, float* dbeta, uint32_t* sync, void* scratch, int64_t scratch_bytes,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * (iv) input / output layers -- scn.InputLayer features (mode 3 sums) and scn.OutputLayer
 * (uresnet_sparse.py:20,24), forward and backward.
 * ------------------------------------------------------------------------------------- */
/* Batch collation (replaces the host-side padding of /root/reference/uresnet/iotools/iotools_sparse.py:131-141):
 * events back to back as int32 coords [N, D] (row stride ldc) + offsets_dev[nb + 1] (event e owns rows
 * [offsets[e], offsets[e+1])) -> out [N, D + 1] int32 (row stride ldo) with the sample id first_id + e in the
 * last column; N = offsets[nb] <= n_cap is also stored to n_out_dev (optional), ready to be the n_dev of the
 * input layer. */
int b200scn_collate(const int32_t* coords, int64_t ldc, const int32_t* offsets_dev, int nb, int D, int first_id,
                    int64_t n_cap, int32_t* out, int64_t ldo, int32_t* n_out_dev, void* stream);

/* dst[i, :] = src[idx[i], :]                        (OutputLayer fwd, InputLayer bwd) */
int b200scn_gather_rows(const float* src, int64_t lds, const int32_t* idx, int64_t n_cap, const int32_t* n_dev,
                        int C, float* dst, int64_t ldd, void* stream);
/* dst[r, :] = src[first_idx[r], :] + sum_{j in seg[seg_off[r] .. seg_off[r+1])} src[j, :]   in that order
 *                                                   (InputLayer fwd, OutputLayer bwd) */
int b200scn_segsum_rows(const float* src, int64_t lds, const int32_t* first_idx, const int32_t* seg_off,
                        const int32_t* seg, int64_t n_cap, const int32_t* n_dev, int C, float* dst, int64_t ldd,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * (v) elementwise helpers used by the table modules (AddTable / JoinTable) and the trainer.
 * ------------------------------------------------------------------------------------- */
/* ---------------------------------------------------------------------------------------
 * (vi) dense path -- the Conv{2,3}d / ConvTranspose{2,3}d / replicate-pad family of
 * /root/reference/uresnet/models/uresnet_dense.py:12-226 on channels-last [B, (D,) H, W, C] fp32.
 * geom_host = int32[11] on the HOST: {mode, dims, in_z, in_y, in_x, out_z, out_y, out_x, k, stride, pad}
 * (z sizes = 1 for 2D); mode 1: in = out*stride + tap - pad (nn.Conv forward, ConvTranspose dgrad),
 * mode 2: in = (out + pad - tap)/stride (nn.ConvTranspose forward, Conv dgrad).  Tap index
 * enumerates the k^dims box with the last dimension fastest; weights via element strides.
 * ------------------------------------------------------------------------------------- */
int b200scn_dense_conv(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                       const float* bias, const int32_t* geom_host, int batch, int Cin, int Cout, float* out,
                       int64_t ldo, void* stream);
/* the same convolution on the tcgen05 tensor cores (implicit GEMM, kind::tf32, accumulators in TMEM;
 * csrc/dense_tc.cu), fp32 accumulate; math_mode B200SCN_MATH_TF32 (single pass) or B200SCN_MATH_TF32X3
 * (3-term split, fp32-equivalent).  Needs Cin % 32 == 0 or Cin == 16, and Cout % 16 == 0, else returns
 * B200SCN_ERR_UNSUPPORTED and the caller uses b200scn_dense_conv. */
int64_t b200scn_dense_conv_tc_scratch_bytes(int taps, int Cin, int Cout);
int b200scn_dense_conv_tc(const float* in, int64_t ldi, const float* W, int64_t w_sk, int64_t w_si, int64_t w_so,
                          const float* bias, const int32_t* geom_host, int batch, int Cin, int Cout, float* out,
                          int64_t ldo, int math_mode, void* scratch, int64_t scratch_bytes, void* stream);
int64_t b200scn_dense_wgrad_scratch_bytes(const int32_t* geom_host, int batch, int Ca, int Cb);
/* dW[k][ca][cb] = sum_o A[idx(o,k)][ca] * B[o][cb] (gather_b == 0) or A[o][ca] * B[idx(o,k)][cb] (gather_b == 1) */
/* math_mode != B200SCN_MATH_FP32 and channel counts that are multiples of 16: tensor-core kernel (csrc/wgrad_mma.cu) */
int b200scn_dense_wgrad(const float* A, int64_t lda, const float* B, int64_t ldb, const int32_t* geom_host, int batch,
                        int gather_b, int Ca, int Cb, float* dW, int math_mode, void* scratch, int64_t scratch_bytes,
                        void* stream);
/* F.pad(mode='replicate') on channels-last data and its backward; sizes are (z, y, x) host int32[3] */
int b200scn_pad_replicate(const float* src, float* dst, int batch, int C, const int32_t* in_sz_host,
                          const int32_t* pl_host, const int32_t* pr_host, int backward, void* stream);
/* out[c] = sum_r x[r, c]  (bias gradients) */
int64_t b200scn_colsum_scratch_bytes(int C);
int b200scn_colsum(const float* x, int64_t ldx, int64_t n, int C, float* out, void* scratch, int64_t scratch_bytes,
                   void* stream);

/* ---------------------------------------------------------------------------------------
 * (vii) fused segmentation head + loss: the Linear(m -> nc) of
 * /root/reference/uresnet/models/uresnet_sparse.py:25,36 and SegmentationLoss.forward (:44-82) --
 * per-event mean cross-entropy (optionally pixel-weighted) summed over events, per-event accuracy summed
 * over events -- plus the whole backward of that chain for an upstream gradient of 1.
 * feat [n, m] = OutputLayer rows; W [nc, m], bias [nc] (torch.nn.Linear layout); label [n] (float class ids,
 * as the reference ships them); ids = the sample-id column (dtype code as for b200scn_input_build, element
 * stride ld_ids); nb = number of events a sample id may name.
 * Outputs: loss_acc[4] = {total_loss, total_acc, rows whose sample id is outside [0, nb), rows whose label is
 * outside [0, nc)} -- such rows contribute nothing (the reference would fail on them, uresnet_sparse.py:57-82);
 * the two counts let the caller refuse the step without a second pass.  Optional logits [n, nc]; optional
 * (training) dfeat [n, m], dW [nc, m], db [nc]; accumulate != 0 ADDS dW / db to what the buffers hold (a
 * persistent flat gradient buffer).  Deterministic (integer counting pass, fixed-order partial sums). */
int64_t b200scn_head_loss_scratch_bytes(int64_t n_cap, int m, int nc, int nb);
int b200scn_head_loss(const float* feat, int64_t ldf, const float* W, const float* bias, const float* label,
                      int64_t ld_label, const float* pix_w, int64_t ld_pw, const void* ids, int id_dtype,
                      int64_t ld_ids, int64_t n_cap, const int32_t* n_dev, int m, int nc, int nb, float* logits,
                      int64_t ld_logits, float* loss_acc, float* dfeat, int64_t ld_dfeat, float* dW, float* db,
                      int accumulate, void* scratch, int64_t scratch_bytes, void* stream);

/* Inference outputs: what /root/reference/uresnet/trainval.py:127-130 copies to the host
 * after a forward pass -- logits ('segmentation') and softmax -- plus the argmax prediction and, when `label`
 * and `metrics` are given, the per-event sums behind the accuracy / confusion part of compute_metrics_sparse
 * (/root/reference/uresnet/utils.py:127-200).  metrics[nb][E], E = 4 + nc + 2 nc nc, per event:
 *   [0] pixels  [1] correctly predicted pixels  [2] sum of -log softmax[label]  [3] sum of softmax[label]
 *   [4 + c] sum of softmax[label] over pixels of true class c
 *   [4 + nc + c nc + c2] pixels of true class c predicted as c2      (confusion_matrix)
 *   [4 + nc + nc nc + c nc + c2] summed `energy` of those pixels     (energy_confusion_matrix; 0 without energy)
 * Rows with a sample id outside [0, nb) or a label outside [0, nc) are skipped by the metrics.  Every output
 * pointer is optional.  ids / energy: dtype codes as for b200scn_input_build, element strides.  nb * E <= 8192. */
int64_t b200scn_head_predict_scratch_bytes(int64_t n_cap, int nc, int nb);
int b200scn_head_predict(const float* feat, int64_t ldf, const float* W, const float* bias, const float* label,
                         int64_t ld_label, const void* ids, int id_dtype, int64_t ld_ids, const void* energy,
                         int energy_dtype, int64_t ld_energy, int64_t n_cap, const int32_t* n_dev, int m, int nc,
                         int nb, float* logits, int64_t ld_logits, float* softmax, int64_t ld_softmax,
                         int32_t* pred, double* metrics, void* scratch, int64_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * (viii) optimiser: torch.optim.Adam (defaults; /root/reference/uresnet/trainval.py:21-30,159-167) over all
 * parameters in one launch.  jobs_dev: n_jobs records of int64[5] = {param, grad, exp_avg, exp_avg_sq, numel};
 * step_dev[1]: number of steps taken so far (incremented by the kernel); ticket_dev[1]: zero-initialised. */
int b200scn_adam_step(const int64_t* jobs_dev, int n_jobs, int64_t max_elems, int32_t* step_dev, uint32_t* ticket_dev,
                      float lr, float beta1, float beta2, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SCN_H */
