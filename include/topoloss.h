/*
 * topoloss.h -- C ABI of libtopoloss.so: the B200 (sm_100a) topological-loss hot path.
 *
 * This is the drop-in boundary for the one path the library replaces:
 *
 *   topo_loss(pred_obj, true_obj, lamda, interp, feat_d, loss_q, loss_r)
 *       /root/reference/octsam/models/topological_loss.py:11-96
 *   called from the SAM fine-tuning step at
 *       /root/reference/octsam/models/training_utils.py:64 (train) and :375 (validation)
 *   and differentiated by  train_loss.backward()  (training_utils.py:66).
 *
 * The reference has no FFI of its own (it is a Python callable over torch_topological ->
 * gudhi / POT); the entry points below are what a ctypes binding of that callable binds:
 * INTEGRATION.md shows the stub.  Plain pointers and sizes only, no torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its comment says host;
 *   - the caller owns every buffer, including the workspace (size from
 *     tl_workspace_bytes); the library never allocates or frees device memory and keeps
 *     no global device state;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises the host with the device;
 *   - every function returns TL_OK (0) or a negative TL_ERR_* code; tl_last_error()
 *     returns a thread-local, human-readable message for the last failure;
 *   - maps are [B, C, H, W] fp32, contiguous (NCHW): one (image, class) map is one
 *     contiguous H*W segment.  Pixels are the top-dimensional cells of the cubical
 *     complex (gudhi T-construction), sublevel filtration, as CubicalComplex(dim=2,
 *     superlevel=False) at topological_loss.py:55-58.
 *   - persistence pairs are (creator pixel, destroyer pixel) flat C-order indices r*W+c.
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef TOPOLOSS_H_
#define TOPOLOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TL_OK 0
#define TL_ERR_ARG (-1)       /* bad argument (shape, feat_d, null pointer, ...) */
#define TL_ERR_WORKSPACE (-2) /* workspace too small */
#define TL_ERR_CUDA (-3)      /* a CUDA runtime call failed; see tl_last_error() */

#define TL_ABI_VERSION 1

/* ABI version of the loaded library (TL_ABI_VERSION it was built with). */
int tl_version(void);

/* Thread-local message for the last error returned on this thread ("" if none). */
const char* tl_last_error(void);

/*
 * Bytes of device workspace needed by tl_forward / tl_backward / tl_persistence_pairs for
 * maps of this shape.  feat_d in {0, 1} is the homology dimension that will be used
 * (batch_iter(..., dim=feat_d), topological_loss.py:68-75).  *bytes is a host pointer.
 */
int tl_workspace_bytes(int B, int C, int H, int W, int feat_d, size_t* bytes);

/*
 * Forward pass of topo_loss for interp == 0 (topological_loss.py:55-96):
 *   per (b, c) map: cubical persistence pairs of pred and truth in dimension feat_d
 *                   (CubicalComplex.forward, :62-63; essential H0 class paired with argmax),
 *   per map:        exact q-Wasserstein matching cost with L-inf ground metric
 *                   (WassersteinDistance(q=loss_q), :78-82),
 *   per image:      W_b = (sum_c cost_{b,c})^(1/q),
 *   loss = lamda / B_global * sum_b W_b                       (:85, :96)
 *          [+ lamda / (B_global*C) * sum_{b,c} sum_pairs |d-b|^q   if loss_r (:88-94)].
 * B_global is the batch size the mean runs over; pass B (or 0) on one GPU, the global
 * batch when the batch axis is sharded over ranks (the caller then all-reduces loss_out).
 * loss_out: one fp32 on the device.  The workspace keeps what tl_backward needs.
 */
int tl_forward(const float* pred, const float* truth, int B, int C, int H, int W,
               int feat_d, float q, float lamda, int loss_r, int B_global,
               void* ws, size_t ws_bytes, float* loss_out, void* stream);

/*
 * Backward pass (the autograd graph of training_utils.py:66 restricted to this loss):
 * grad_pred[B,C,H,W] = grad_loss * d loss / d pred, fully overwritten (zeros included).
 * grad_loss: one fp32 on the device (upstream gradient), or NULL for 1.0.
 * `ws` must be the workspace a tl_forward call filled; shape, feat_d, q, lamda, loss_r and
 * B_global must be the values given to that call (the library keeps no state between calls).
 * An image whose summed cost S_b is exactly 0 gets NaN on all its critical pixels, as the
 * reference's autograd does (0 * inf through pow(1/q)).
 */
int tl_backward(const float* grad_loss, const void* ws, size_t ws_bytes,
                int B, int C, int H, int W, int feat_d, float q, float lamda, int loss_r,
                int B_global, float* grad_pred, void* stream);

/*
 * Inner boundary for parity tests: CubicalComplex.forward on n_maps independent HxW maps
 * (torch_topological CubicalComplex._forward -> gudhi persistence +
 * cofaces_of_persistence_pairs).  Writes, per map, up to `cap` pairs
 * (creator, destroyer) of homology dimension `dim` into pairs[map][k][2], sorted in
 * gudhi's emission order (filtration order of the death cell; essential H0 class last),
 * and the count into counts[map].  A map with more than `cap` pairs reports its true count
 * and writes only the first `cap`.
 */
int tl_persistence_pairs(const float* maps, int n_maps, int H, int W, int dim,
                         void* ws, size_t ws_bytes,
                         int32_t* pairs, int cap, int32_t* counts, void* stream);

/* Largest number of pairs one HxW map can produce in dimension dim (buffer sizing). */
int tl_max_pairs(int H, int W, int dim);

/*
 * Inner boundary for parity tests: WassersteinDistance cost of one channel, batched.
 * Diagram k of set 1 is rows off1[k]..off1[k+1]-1 of D1 ([rows][2] fp32 (birth, death)),
 * likewise D2/off2 (off arrays have n_diag+1 int32 entries, device).  Writes
 * cost[k] (fp64, the emd2 value before the 1/q root) and match1[row] = row index inside
 * diagram k of D2 that D1's row is matched to, or -1 for the diagonal.
 */
int tl_wasserstein(const float* D1, const int32_t* off1, const float* D2, const int32_t* off2,
                   int n_diag, int max_rows1, int max_rows2, float q,
                   void* ws, size_t ws_bytes, double* cost, int32_t* match1, void* stream);

/*
 * Debug aid, not on the hot path: host copy of the 8 per-phase cycle counters the persistence
 * kernel accumulates into the workspace when the process environment has TL_PROFILE=1.
 * Synchronises the device.  host_out8: 8 x uint64 on the host.
 */
int tl_debug_profile(const void* ws, unsigned long long* host_out8);

/*
 * Measurement aid for bench.py's roofline: while enabled on the calling thread, tl_forward and
 * tl_backward bracket each of their kernels with cudaEventRecord on the caller's stream (the
 * stream the kernels are launched on).  tl_timing_read synchronises those events and returns, on
 * the host, the summed milliseconds of the 6 stages [persistence, segmented sort, matching, loss,
 * grad zero-fill, grad scatter] over the calls since the last enable/read (at most 128), and the
 * number of forward / backward calls in n_calls[0..1].
 */
int tl_timing_enable(int on);
int tl_timing_read(float* ms_sum6, int* n_calls);

/*
 * F1, the step in front of the path at the reference call site
 *   topo_loss(torch.sigmoid(masks.float()), gt_masks.float(), 0.1, feat_d=1, interp=50)
 *       /root/reference/octsam/models/training_utils.py:64
 * i.e. torch.sigmoid followed by F.interpolate(size=(interp, interp), mode="bilinear",
 * align_corners=True) of /root/reference/octsam/models/topological_loss.py:33-46, fused:
 *   out[m, oy, ox] = bilinear_{align_corners}( apply_sigmoid ? sigmoid(in[m]) : in[m] )(oy, ox)
 * in: [n_maps][H][W] fp32 (logits when apply_sigmoid, e.g. the mask decoder output; plain maps
 * otherwise, e.g. ground truth), out: [n_maps][S][S].  Only the <= 4 S^2 source pixels an output
 * needs are read and passed through the sigmoid.
 */
int tl_resample_forward(const float* in, int n_maps, int H, int W, int S, int apply_sigmoid,
                        float* out, void* stream);

/*
 * Backward of tl_resample_forward: grad_in[n_maps][H][W] (16-byte aligned) is fully overwritten
 * with sum over outputs of weight * grad_out * (apply_sigmoid ? s(1-s) : 1); `in` are the forward inputs.
 */
int tl_resample_backward(const float* grad_out, const float* in, int n_maps, int H, int W, int S,
                         int apply_sigmoid, float* grad_in, void* stream);

/*
 * F3: SAM post-processing of the mask decoder output, fused into one gather:
 *   m   = F.interpolate(pred_masks.squeeze(2), (T, T), mode="bilinear", align_corners=False)   T = 1024
 *   m   = m[..., :rh, :rw]                                                  (reshaped_input_sizes)
 *   out = F.interpolate(m, (oh, ow), mode="bilinear", align_corners=False)  (original_sizes)
 *       /root/reference/octsam/models/training_utils.py:57-59
 * in: [n_maps][Hs][Ws] fp32, out: [n_maps][oh][ow].  The T x T intermediate is never materialised.
 */
int tl_postprocess_forward(const float* in, int n_maps, int Hs, int Ws, int T, int rh, int rw,
                           int oh, int ow, float* out, void* stream);

/* Backward of tl_postprocess_forward: grad_in[n_maps][Hs][Ws] (16-byte aligned) fully overwritten. */
int tl_postprocess_backward(const float* grad_out, int n_maps, int Hs, int Ws, int T, int rh, int rw,
                            int oh, int ow, float* grad_in, void* stream);

/* Bytes of workspace tl_wasserstein needs. */
int tl_wasserstein_workspace_bytes(int n_diag, int max_rows1, int max_rows2, size_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* TOPOLOSS_H_ */
