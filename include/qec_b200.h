/* qec_b200.h -- C ABI of the B200-native QEC-Net quaternion dynamic-routing hot path.
 *
 * The reference (tolgabirdal/qecnetworks) has no C ABI: its native code is two
 * pybind11 modules (models/pytorch-cusolver/torch_cusolver.cpp:558-560,
 * models/pytorch-autograd-solver/torch_autograd_solver.cpp:267-273) and the rest
 * of the hot path is ATen ops issued from models/qec_module.py.  These entry
 * points are what a binding for that path would call; every one cites the
 * reference interface it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous FP32 (int64 / int32 where
 *    stated) on the current device; quaternions are (w,x,y,z), 16-byte aligned;
 *  - the caller owns and allocates every buffer; the library never allocates and
 *    never synchronises.  Its only process-wide state is launch configuration
 *    (occupancy / the shared-memory opt-in of the cluster kernels), cached PER
 *    DEVICE ORDINAL in atomics, and the build selector of
 *    qec_routing_fused_set_wide (an atomic), so it is re-entrant: one thread per
 *    GPU under the reference's nn.DataParallel (the calling thread's current
 *    device must be the device of the pointers), one process per GPU under DDP;
 *  - `stream` is a cudaStream_t (NULL = legacy default stream); all work is
 *    enqueued on it and is CUDA-graph capturable;
 *  - return value: 0 on success; -k if the k-th argument is invalid (mirrors
 *    check_jacobi's "k-th parameter is wrong", torch_cusolver.cpp:162-166);
 *    a positive cudaError_t if the launch failed.
 *
 * Shapes: B batch, P patches, K neighbours, I in_channels, O out_channels,
 * T routing iterations, BP = B*P, J = K*I votes per output capsule.
 */
#ifndef QEC_B200_H
#define QEC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* Library/ABI version (major*100 + minor). */
int qec_b200_version(void);

/* mean_lrf_per_patch + averageQuaternions + the qrotv of get_votes
 * (models/qec_module.py:101-110, 35-61, 115-120; models/quat_ops.py:50-68).
 *   lrfs [BP,K,I,4], act [BP,K,I], points [BP,K,3]
 *   -> mean [BP,I,4], xrot [BP,K,I,3] (= the input of quater_gen1, i-major xyz-minor).
 * xrot (and then points) may be NULL to skip the rotation. */
int qec_lrf_mean_fwd(const float* lrfs, const float* act, const float* points, float* mean, float* xrot, int BP,
                     int K, int I, void* stream);

/* Backward of the above (autograd through S.symeig in averageQuaternions:
 * torch_autograd_solver/__init__.py:83-98 -> torch_autograd_solver.cpp:100-169).
 *   g_xrot [BP,K,I,3], g_mean [BP,I,4] (nullable) -> g_lrfs [BP,K,I,4] (overwritten),
 *   g_points [BP,K,3] (nullable; ACCUMULATED, caller zero-fills). */
int qec_lrf_mean_bwd(const float* lrfs, const float* points, const float* mean, const float* g_xrot,
                     const float* g_mean, float* g_lrfs, float* g_points, int BP, int K, int I, void* stream);

/* get_votes after the two Linears (models/qec_module.py:126-147; quat_ops.qmul :30-48).
 *   lrfs [BP,K,I,4], t_raw [BP*K, 4*I*O] in the reference feature order q*I*O + i*O + o
 *   -> votes [BP,O,J,4]. */
int qec_votes_fwd(const float* lrfs, const float* t_raw, float* votes, int BP, int K, int I, int O, void* stream);

/* Backward: g_votes [BP,O,J,4] -> g_traw [BP*K,4*I*O] (overwritten), g_lrfs [BP,K,I,4]
 * (nullable; overwritten if O <= 64, otherwise ACCUMULATED: caller zero-fills). */
int qec_votes_bwd(const float* lrfs, const float* t_raw, const float* g_votes, float* g_traw, float* g_lrfs,
                  int BP, int K, int I, int O, void* stream);

/* The routing loop + activation of QecModule.forward over materialised votes
 * (models/qec_module.py:172-194 with mean/weightedAverageQuaternions :149-153, 64-98,
 * distance :156-160; the S.symeig call at :90 is the in-kernel Jacobi).
 *   votes [BP,O,J,4], act [BP,J], alpha [1], beta [1]
 *   -> pose [BP,O,4], agreement [BP,O],
 *      poses_all [T,BP,O,4] and stats [BP,O,2] (mean inlier score, active-vote count):
 *      saved for the backward.  1 <= T <= 10. */
int qec_routing_fwd(const float* votes, const float* act, const float* alpha, const float* beta, float* pose,
                    float* agreement, float* poses_all, float* stats, int BP, int O, int J, int T, void* stream);

/* Backward through all T iterations (eigenvector adjoint in closed form).
 *   g_pose [BP,O,4], g_agreement [BP,O]
 *   -> g_votes [BP,O,J,4] (overwritten),
 *      g_act [BP,J] (nullable = activation needs no gradient, only the last iteration is
 *      differentiated; otherwise ACCUMULATED, caller zero-fills),
 *      g_alpha_beta [2] (ACCUMULATED, caller zero-fills). */
int qec_routing_bwd(const float* votes, const float* act, const float* alpha, const float* beta,
                    const float* poses_all, const float* stats, const float* g_pose, const float* g_agreement,
                    float* g_votes, float* g_act, float* g_alpha_beta, int BP, int O, int J, int T, void* stream);

/* Fused, cluster-resident variant of qec_votes_* + qec_routing_* for capsules with thousands
 * of votes (J = K*I up to 32768; the pool2 layer).  Same reference functions as above
 * (models/qec_module.py:126-147 and 172-194) but the [BP,O,J,4] vote tensor is never
 * materialised: votes are generated into shared memory across a thread-block cluster.
 *   t_oqi [BP*K, 4*I*O]: the output of quater_gen2 with its weight rows permuted to the
 *   capsule-major feature order o*4I + q*I + i (instead of q*I*O + i*O + o).
 * qec_routing_fused_supported returns 1/0 (not a status).  Backward: g_toqi [BP*K,4*I*O]
 * overwritten; g_lrfs [BP,K,I,4] and g_act [BP,J] nullable and ACCUMULATED (caller zero-fills);
 * g_alpha_beta [2] ACCUMULATED. */
int qec_routing_fused_supported(int K, int I, int O, int T);
int qec_routing_fused_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                          const float* beta, float* pose, float* agreement, float* poses_all, float* stats, int BP,
                          int K, int I, int O, int T, void* stream);
int qec_routing_fused_bwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                          const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                          const float* g_agreement, float* g_toqi, float* g_lrfs, float* g_act, float* g_alpha_beta,
                          int BP, int K, int I, int O, int T, void* stream);

/* Composed-weight variant for in_channels = 1 (the pool1 layer; BASELINE config 5's stress geometry): the
 * two Linears of get_votes have no non-linearity between them (models/qec_module.py:123-124), so
 * t_raw = Wc x + bc with Wc = W2 W1 [4*O,3], bc = W2 b1 + b2 [4*O] (row q*O + o) composed by the
 * caller; the kernel covers qec_module.py:123-147 and 172-194 without the [BP*K, 4*O] tensor.
 *   xrot [BP,K,3] (from qec_lrf_mean_fwd), lrfs [BP,K,4], act [BP,K]
 * qec_routing_small_supported: 1 = K <= 16, votes register-resident, forward and backward; 2 = 16 < K <= 64,
 * forward only, votes regenerated in every sweep (nothing of the size of the vote tensor exists); 0 = not served.
 * poses_all [T,BP,O,4] and stats [BP,O,2] are what the backward reads: both may be NULL in the forward (inference).
 * Backward: only parameter gradients (g_Wc [4*O,3], g_bc [4*O], g_alpha_beta [2], all ACCUMULATED);
 * callers that need dL/dlrfs, dL/dact or dL/dpoints use qec_votes_* + qec_routing_*. */
int qec_routing_small_supported(int K, int I, int O, int T);
int qec_routing_small_fwd(const float* xrot, const float* lrfs, const float* act, const float* Wc, const float* bc,
                          const float* alpha, const float* beta, float* pose, float* agreement, float* poses_all,
                          float* stats, int BP, int K, int O, int T, void* stream);
int qec_routing_small_bwd(const float* xrot, const float* lrfs, const float* act, const float* Wc, const float* bc,
                          const float* alpha, const float* beta, const float* poses_all, const float* stats,
                          const float* g_pose, const float* g_agreement, float* g_Wc, float* g_bc,
                          float* g_alpha_beta, int BP, int K, int O, int T, void* stream);

/* torch_autograd_solver.symeig on [n,4,4] (torch_autograd_solver/__init__.py:141-159,
 * 74-81 -> torch_cusolver.cpp:194-288): w [n,4] ascending, V [n,4,4] with ROWS =
 * eigenvectors, upper triangle of the row-major input.
 * status (nullable): device int, incremented once per matrix that did not converge
 * (replaces check_jacobi's blocking host loop, torch_cusolver.cpp:150-173). */
int qec_symeig4_fwd(const float* M, float* w, float* V, int* status, long long n, void* stream);

/* batch_symeig_backward (torch_autograd_solver.cpp:100-169): gw [n,4] and/or gV [n,4,4]
 * (rows convention; either may be NULL) -> gM [n,4,4].  fold != 0 applies the
 * reference's final triu(g) + triu(g^T,1). */
int qec_symeig4_bwd(const float* gw, const float* gV, const float* w, const float* V, float* gM, long long n,
                    int fold, void* stream);

/* Loader neighbour gather (my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160).
 *   point_set [B,N,3], lrf_set [B,N,4], idx [B,P,K] int64 (-1 = empty),
 *   wrong_ids [B,Wmax] int64 (nullable), wrong_count [B] int32 (nullable = Wmax each)
 *   -> points [B,P,K,3], lrfs [B,P,K,4], act [B,P,K];
 *   err: device int set to 1 if an index is out of range (reference: IndexError). */
int qec_gather(const float* point_set, const float* lrf_set, const long long* idx, const long long* wrong_ids,
               const int* wrong_count, float* points, float* lrfs, float* act, int* err, int B, int N, int P, int K,
               int Wmax, void* stream);

/* The two generations behind qec_routing_fused_*: the packed-pair kernels (FFMA2, even I;
 * routing_fused2.cu) are used whenever qec_routing_fused_packed_supported returns 1, otherwise
 * the scalar kernels, which are also exported on their own so that tests can cross-check the
 * two on the same inputs.  Same arguments and semantics as qec_routing_fused_fwd / _bwd. */
int qec_routing_fused_packed_supported(int K, int I, int O, int T);
int qec_routing_fused_fwd_scalar(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                 const float* beta, float* pose, float* agreement, float* poses_all, float* stats,
                                 int BP, int K, int I, int O, int T, void* stream);
int qec_routing_fused_bwd_scalar(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                 const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                                 const float* g_agreement, float* g_toqi, float* g_lrfs, float* g_act,
                                 float* g_alpha_beta, int BP, int K, int I, int O, int T, void* stream);

/* qec_routing_fused_bwd with dL/dt_oqi delivered pre-split for tensor-core consumers (the backward
 * GEMMs of quater_gen2 at FP32 accuracy): g_toqi_hi = the gradient rounded to the TF32 grid (FP32
 * storage), g_toqi_lo = the remainder g - hi as BF16 (same shape; |g - hi - lo| <= 2^-20 |g|).  Only where
 * qec_routing_fused_packed_supported(...) == 1; returns -15 otherwise.  No reference counterpart
 * (the reference's backward of nn.Linear is an FP32 cuBLAS SGEMM, models/qec_module.py:124). */
int qec_routing_fused_bwd_split(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                                const float* beta, const float* poses_all, const float* stats, const float* g_pose,
                                const float* g_agreement, float* g_toqi_hi, unsigned short* g_toqi_lo, float* g_lrfs,
                                float* g_act, float* g_alpha_beta, int BP, int K, int I, int O, int T, void* stream);

/* Deterministic K-nearest-neighbour indices of the patch centres (on-device
 * counterpart of the neighbourhood index files the reference's loader consumes,
 * my_dataloader/modelnet_with_lrf_sample_index_loader.py:130-160, produced offline by
 * my_dataloader/gen_downsample_m40.py:86-150).  d = ((cx-x)^2 + (cy-y)^2) + (cz-z)^2 with every
 * operation rounded to float32 separately; neighbours sorted by (d, index) ascending -> bit-exact
 * against the float32 CPU statement oracle/qec_oracle.py:knn_indices.  points [B,N,3]; centres given
 * either as point indices (centre_idx [B,P] int64) or as coordinates (centres [B,P,3]), exactly one
 * non-null; idx_out [B,P,K] int64 (-1 where N < K); K <= 64.  err: device int set to 1 if a centre
 * index is outside [0, N).  The output feeds qec_gather. */
int qec_knn(const float* points, const float* centres, const long long* centre_idx, long long* idx_out, int* err,
            int B, int N, int P, int K, void* stream);

/* Operand preparation for the quater_gen2 GEMMs on the tensor cores at FP32 accuracy ("3xTF32";
 * reference: nn.Linear in models/qec_module.py:124 and its autograd backward, FP32 cuBLAS SGEMM).
 * qec_split_operand: src [R, C] (+ one extra column when add_col: extra[R] if given, else
 *   extra_const -- the bias / the column of ones that folds the bias into the GEMM) = Kc real columns.
 *   out_f32 [R, WT] (nullable): nblk <= 3 blocks at stride BS >= Kc, block j = the TF32 hi part
 *   (blkj == 1) or the lo part (blkj == 2), zeros elsewhere; out_bf16 [R, WB] (nullable): the hi part in
 *   BF16, zero-padded.  Output row r is read from source row perm[r] when perm (int64 [R]) is given.
 * qec_fold_chunks: a [S, R, 2*KP], b [S, R, KP] split-K partial products -> out_main [*, ld_main] and
 *   (ld_main == K - 1) column K - 1 -> out_last [*] (nullable = dropped); row r lands in row perm[r]. */
int qec_split_operand(const float* src, const float* extra, float extra_const, const long long* perm,
                      float* out_f32, unsigned short* out_bf16, long long R, int C, int add_col, int BS, int WT,
                      int WB, int blk0, int blk1, int blk2, int nblk, void* stream);
int qec_fold_chunks(const float* a, const float* b, const long long* perm, float* out_main, float* out_last, int S,
                    long long R, int KP, int K, int ld_main, void* stream);

/* Backward of the transform generator quater_gen2 = nn.Linear(O, 4*I*O) of the pool2 layer
 * (models/qec_module.py:124 and its autograd) on the tcgen05 tensor cores at FP32 accuracy (3-term TF32 split,
 * accumulator in tensor memory; csrc/gen2_bwd.cu):
 *   g [M, N] = dL/dt (FP32, as qec_routing_fused_bwd writes it; M = B*P*K, N = 4*I*O in capsule-major column order),
 *   h [M, Kc] = the layer's input, W2 [N, Kc] = its weight in the reference's row order, perm [N] int64 = the row of
 *   W2 behind column n of g  ->  g_h [M, Kc] = g Wp (NULL to skip), g_W2 [N, Kc] and g_b2 [N] (both or neither).
 * workspace: qec_gen2_bwd_workspace(M, N, Kc) floats.  Supported: M, N multiples of 128, Kc <= 47
 * (qec_gen2_bwd_supported).  variant: 0 (other values select descriptor-stride conventions; bring-up aid). */
int qec_gen2_bwd_supported(int M, int N, int Kc);
long long qec_gen2_bwd_workspace(int M, int N, int Kc);
int qec_gen2_bwd(const float* g, const float* h, const float* W2, const long long* perm, float* g_h, float* g_W2,
                 float* g_b2, float* workspace, int M, int N, int Kc, int variant, void* stream);

/* ---- forward of quater_gen2 on the tcgen05 tensor cores ---------------------------------------------------------
 * reference models/qec_module.py:124 (nn.Linear(O, 4*I*O) forward, bias included), output columns permuted:
 *   t [M, N] = h [M, Kc] W2[perm]^T + b2[perm]     (W2 [N, Kc], b2 [N] in the reference's row order; perm [N] int64
 *   | NULL = row of W2 / b2 behind output column n).  TMA-fed, 3-term TF32 split at FP32 accuracy, accumulators in
 * tensor memory, persistent CTAs.  workspace: qec_gen2_fwd_workspace(M, N, Kc) floats.
 * Supported (qec_gen2_fwd_supported): M, N multiples of 128, Kc <= 63. */
int qec_gen2_fwd_supported(int M, int N, int Kc);
long long qec_gen2_fwd_workspace(int M, int N, int Kc);
int qec_gen2_fwd(const float* h, const float* W2, const float* b2, const long long* perm, float* t, float* workspace,
                 int M, int N, int Kc, void* stream);

/* ---- narrow FP32 Linear (out features O <= 64) and the composition of the two Linears of get_votes ----------------
 * reference models/qec_module.py:123 (quater_gen1 = nn.Linear(3*I, O), forward and autograd backward) and :123-124
 * (quater_gen2(quater_gen1(x)) with no non-linearity in between: for I = 1 the register-resident kernels evaluate the
 * composed map Wc x + bc).  Plain FP32 FMA, fixed summation order.
 *   qec_linear_fwd:  y [M, O] = x [M, C] W^T + b                       (W [O, C], b [O])
 *   qec_linear_bwd:  g [M, O] -> g_x [M, C] | NULL, g_W [O, C] | NULL, g_b [O] | NULL;
 *                    workspace: qec_linear_bwd_workspace(M, C, O) floats
 *   qec_compose_fwd: Wc [R, C] = W2 [R, O] W1 [O, C],  bc [R] = W2 b1 + b2        (C <= 16: qec_compose_supported)
 *   qec_compose_bwd: (g_Wc, g_bc) -> g_W1 [O, C], g_b1 [O], g_W2 [R, O], g_b2 [R] | NULL  (dL/db2 = g_bc, copied) */
int qec_linear_supported(int C, int O);
long long qec_linear_bwd_workspace(int M, int C, int O);
int qec_linear_fwd(const float* x, const float* W, const float* b, float* y, int M, int C, int O, void* stream);
int qec_linear_bwd(const float* x, const float* W, const float* g, float* g_x, float* g_W, float* g_b, float* workspace,
                   int M, int C, int O, void* stream);
int qec_compose_supported(int C);
int qec_compose_fwd(const float* W1, const float* b1, const float* W2, const float* b2, float* Wc, float* bc, int R, int O,
                    int C, void* stream);
int qec_compose_bwd(const float* W1, const float* b1, const float* W2, const float* g_Wc, const float* g_bc, float* g_W1,
                    float* g_b1, float* g_W2, float* g_b2, int R, int O, int C, void* stream);

/* Spread loss of QEC-Net and its gradient in one launch pair (reference models/qec_net.py:53-59,
 * margin 0.1): loss [1] = mean_b sum_{c != y_b} relu(m - (x[b,y_b] - x[b,c]))^2, g_x [B, C] = dloss/dx.
 * scratch: B floats.  err: device int set to 1 if a label is outside [0, C) (the reference's
 * scatter_ raises). */
int qec_spread_loss(const float* x, const long long* labels, float margin, float* loss, float* g_x, float* scratch,
                    int* err, int B, int C, void* stream);

/* The losses of the two networks with their gradients in ONE launch: spread loss of S = 1
 * (models/qec_net.py:53-59) or S = 2 branches (models/qec_sia_net.py:69-77) plus, if pose != NULL, w_pose times the
 * pose-difference loss (qec_sia_net.py:79-82) on the poses of the ground-truth class, i.e. including the per-sample
 * selection loop of train_cls_sia.py:72-78 (pose_out_act[:, i] = pose_out[:, i, y_i]).
 *   x [S,B,C] (NULL with S = 0: pose term only), labels [B] int64, pose [2,B,C,4] | NULL
 *   -> loss [3] = {total, spread, pose_diff}, g_x [S,B,C], g_pose [2,B,C,4] | NULL (zero off the label class).
 * err: device int set to 1 if a label is outside [0, C). */
int qec_class_pose_loss(const float* x, const long long* labels, float margin, const float* pose, float w_pose,
                        float* loss, float* g_x, float* g_pose, int* err, int S, int B, int C, void* stream);

/* Evaluation-side quaternion ops of test_rot_sia.py, batched: pose1, pose2 [n,4] ->
 * dist [n] | NULL = (2/pi) acos(min(|<fix(p1), fix(p2)>|, 1)) (q_distance, test_rot_sia.py:32-41),
 * rel [n,4] | NULL = conj(pose1) (x) pose2 (the relative pose, test_rot_sia.py:131-132). */
int qec_quat_eval(const float* pose1, const float* pose2, float* dist, float* rel, long long n, void* stream);

/* Which build of the packed cluster-resident routing kernels serves capsules of more than 16 384 votes:
 * 0 = 256 threads per CTA, two CTAs per SM, clusters of 8 (default); 1 = 512 threads, one CTA per SM, clusters
 * of 4, forward and backward; 2 = forward only; 3 = backward only.  Same results either way (tests run both).
 * Initial value from the environment (QEC_FUSED_WIDE = 0 | 1 | f | b).  Returns the previous mode; an
 * out-of-range argument only queries.  No reference counterpart. */
int qec_routing_fused_set_wide(int mode);

/* Which generation of the packed cluster-resident kernels qec_routing_fused_fwd / _bwd dispatch to when both serve
 * the geometry: bit 0 = forward, bit 1 = backward on the third generation (two capsules in flight per cluster, a
 * dedicated solver warp hides the reduce -> solve -> publish chain behind the other capsule's sweep;
 * csrc/routing_fused3.cu), else the second (csrc/routing_fused2.cu).  Same results to rounding (tests run both).
 * Initial value from the environment (QEC_FUSED_PP).  Returns the previous mode; an out-of-range argument only
 * queries.  qec_routing_fused_pp_fwd is the third-generation forward itself (same contract as
 * qec_routing_fused_fwd; argument 11 if qec_routing_fused_pp_supported(...) is 0).  No reference counterpart. */
int qec_routing_fused_set_pp(int mode);
int qec_routing_fused_pp_supported(int K, int I, int O, int T);
int qec_routing_fused_pp_fwd(const float* t_oqi, const float* lrfs, const float* act, const float* alpha,
                             const float* beta, float* pose, float* agreement, float* poses_all, float* stats, int BP,
                             int K, int I, int O, int T, void* stream);

/* One Adam step (reference train_cls.py:44-51: torch.optim.Adam, default betas / eps) on FLAT FP32 buffers of n
 * elements, 16-byte aligned -- parameters, gradient, first and second moment (qecnetworks_b200/ddp.py lays the
 * network's tensors out as views of such buffers).  The gradient is read as grads * grad_scale (1/world after a
 * summing all-reduce) and overwritten with zeros when zero_grad != 0, so the step needs neither a scaling pass
 * nor a memset.  Bias corrections 1 - beta^t are evaluated in double like torch's; weight_decay is torch's L2
 * form (added to the gradient).  state: two device ints, zeroed by the caller before the first step ({steps
 * taken, scratch}); the kernel advances the count itself, so the call is CUDA-graph capturable. */
int qec_adam_flat(float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n, float lr, double beta1,
                  double beta2, float eps, float weight_decay, float grad_scale, int zero_grad, int* state, void* stream);

/* FP32 FMA roof: blocks*256 threads x iters x 16 dependent-chain FMAs
 * (2*16*iters*256*blocks FLOP); timed by the caller with CUDA events.  No reference
 * counterpart: it measures the denominator of the FP32 roofline on the box. */
int qec_ffma_peak(float* out, int blocks, int iters, void* stream);

/* Issue-slot / pipe-sharing probe for the packed FP32 instructions of sm_100 (FFMA2):
 * mode 0: 16 FFMA | 1: 16 FFMA2 | 2: 8 FFMA + 8 FMNMX | 3: 8 FFMA2 + 8 FMNMX | 4: 8 FFMA2 + 8 FFMA |
 * 5: 8 FFMA2 | 6: 8 FFMA | 7: 16 FMNMX chains per thread, iters rounds.  No reference counterpart. */
int qec_pipe_probe(float* out, int blocks, int iters, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QEC_B200_H */
