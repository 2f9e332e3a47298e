/* gaitb200.h - C ABI of the B200-native (sm_100a) regression head for MAX-GRNet.
 *
 * The reference (lisqzqng/Video-based-gait-analysis-for-dementia) is pure Python/PyTorch: it has
 * no FFI, plugin or operator registry.  Its "interface" for the hot path is a set of nn.Module
 * classes and free functions.  This header is the boundary a native backend for
 * those classes binds: every entry point names the reference function (file:line under
 * /root/reference, or the smplx==0.1.26 routine the reference calls) whose arithmetic it replaces.
 * The Python mirror of the reference API (package gaitb200) binds these through ctypes; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: pointers, sizes, a stream handle.  No torch types, no exceptions, no allocation:
 *     the caller passes outputs and workspaces.
 *   - every pointer is a DEVICE pointer to FP32 (or int32 where said), row-major, 16-byte aligned
 *     base unless noted; `ld*` arguments are row strides in elements.
 *   - work is enqueued on `stream` (a cudaStream_t) and the call returns immediately;
 *     calls are re-entrant per stream.
 *   - return 0 on success, a negative GAIT_ERR_* otherwise; gait_last_error() gives detail.
 *   - inputs are never written; in/out aliasing is not allowed unless stated.
 */
#ifndef GAITB200_H
#define GAITB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GAIT_ABI_VERSION 1

#define GAIT_OK 0
#define GAIT_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, bad stride/alignment) */
#define GAIT_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed */
#define GAIT_ERR_UNSUPPORTED (-3) /* shape outside what the kernels are written for */
#define GAIT_ERR_WORKSPACE (-4)   /* workspace too small */

#define GAIT_NUM_JOINTS 24        /* SMPL kinematic tree */
#define GAIT_NUM_BETAS 10
#define GAIT_POSE_BASIS 207       /* 23 * 9 */
#define GAIT_BLEND_K 218          /* 207 pose-blend + 10 shape-blend + 1 (template) */
#define GAIT_BLEND_LD 224         /* row stride of the packed blend operands (K padded, zero filled) */

typedef void* gait_stream_t;      /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------------- */
int gait_abi_version(void);
const char* gait_error_string(int code);
const char* gait_last_error(void);            /* thread-local detail of the last failure */
/* sm count / compute capability of the current device; fails unless it is sm_100. */
int gait_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* A/B hook: sets the mask of kernel kinds launched with programmatic dependent launch (1 tensor-core GEMM, 2 skinning,
 * 4 small kernels; default 3, or GAITB200_PDL) and returns the previous mask; a negative value restores the default.
 * Graphs captured earlier keep the edges they were captured with. */
int gait_debug_pdl_mask(int mask);

/* number of kernels this library has launched in the calling process (bench.py gpu_launches) */
int64_t gait_launch_count(void);

/* ---- geometry: lib/utils/geometry.py ---------------------------------------------------- */
/* rot6d_to_rotmat geometry.py:395-410 (eps=1e-6) and rot6d_to_rotmat_spin :368-387 (eps=1e-12).
 * 6-vector i (interleaved (3,2)) is read at x6 + (i / group) * in_group_stride + (i % group) * 6
 * (group=1, stride=6 for a plain (n,6) array; 24/160 reads the regressor state in place) -> R (n,3,3). */
int gait_rot6d_to_rotmat(const float* x6, int group, int64_t in_group_stride, float* R, int64_t n, float eps,
                         gait_stream_t stream);
/* rotmat_to_rot6d geometry.py:389-393. R (n,3,3) -> (n,6). */
int gait_rotmat_to_rot6d(const float* R, float* x6, int64_t n, gait_stream_t stream);
/* rotation_matrix_to_quaternion geometry.py:213-293. R (n,3,row_stride), row_stride 3 or 4 -> (n,4) wxyz. */
int gait_rotmat_to_quaternion(const float* R, int row_stride, float* quat, int64_t n, float eps,
                              gait_stream_t stream);
/* quaternion_to_angle_axis geometry.py:159-210. (n,4) -> (n,3). */
int gait_quaternion_to_axis_angle(const float* quat, float* aa, int64_t n, gait_stream_t stream);
/* rotation_matrix_to_angle_axis geometry.py:68-97 (quaternion route, NaN -> 0).
 * Rotation i is written to aa[(i / group) * out_group_stride + out_offset + (i % group) * 3]
 * (group=1, out_group_stride=3, out_offset=0 for a plain (n,3) result; 24/85/3 packs theta). */
int gait_rotmat_to_axis_angle(const float* R, int row_stride, float* aa, int64_t n, int group,
                              int64_t out_group_stride, int out_offset, gait_stream_t stream);
/* quat2mat geometry.py:38-65. (n,4) wxyz -> (n,3,3). */
int gait_quat2mat(const float* quat, float* R, int64_t n, gait_stream_t stream);
/* Rodrigues.  variant 0: smplx lbs.batch_rodrigues (I + sin K + (1-cos) K^2), used by
 * SMPL.forward(pose2rot=True) (lib/utils/smooth_pose.py:72-76);  variant 1: geometry.py:23-35
 * (half-angle quaternion -> quat2mat).  aa (n,3) -> R (n,3,3). */
int gait_batch_rodrigues(const float* aa, float* R, int64_t n, int variant, gait_stream_t stream);
/* convert_weak_perspective_to_perspective geometry.py:427-446: cam (n,3)=[s,tx,ty] -> (n,3). */
int gait_weak_perspective_to_translation(const float* cam, float* trans, int64_t n, float focal_length,
                                         float img_res, gait_stream_t stream);
/* perspective_projection geometry.py:448-479.  points (b,j,3); rotation (b,3,3) or NULL (identity);
 * translation (b,3); center (b,2) or NULL (zeros); out (b,j,2) = K[(R p + t)/z] / out_divisor. */
int gait_perspective_projection(const float* points, const float* rotation, const float* translation,
                                const float* center, float focal_length, float out_divisor, float* out,
                                int64_t b, int j, gait_stream_t stream);

/* ---- dense layer: torch.nn.Linear as used at spin.py:216-222,259-265 ---------------------- */
/* C[m,n] = sum_k A[m,k] * W[n,k] + bias[n] + Cin[m,n]   (bias, Cin optional; Cin may alias C).
 * FP32 result: SIMT FP32 or FP32-accurate split-TF32 tensor-core path, chosen by shape. */
int gait_linear(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias,
                const float* Cin, int64_t ldcin, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                gait_stream_t stream);

/* Prepared weights.  The tensor-core path splits every FP32 operand into TF32 hi and lo parts; for a CONSTANT weight matrix
 * this is done once: gait_prepare_weight writes [hi = RN_tf32(W) | lo = RN_tf32(W - hi)] (2n floats, n a multiple of 4) into the
 * caller's buffer W_hilo and registers it, after which every gait_linear / gait_gru_layer / gait_hmr_regressor /
 * gait_smpl_blend call whose weight pointer lies inside [W, W+n) loads both tiles by TMA instead of converting the raw weights
 * in shared memory (same result up to FP32 rounding; the offline split is the round-to-nearest one).  W and W_hilo must stay
 * allocated and unchanged until gait_release_weight(W) (call it before freeing or modifying W). */
int gait_prepare_weight(const float* W, float* W_hilo, int64_t n, gait_stream_t stream);
int gait_release_weight(const float* W);
/* The same with an EXPLICIT handle and no global state: gait_split_weight only writes W_hilo = [hi | lo] (2n floats) for the
 * n floats at W_base; gait_linear_prepared is gait_linear with that operand passed in (W may point anywhere inside
 * [W_base, W_base + n_prepared): row views of a packed weight array).  Nothing is registered, so nothing can go stale: the
 * caller owns W_base / W_hilo and their lifetime.  (The registry above is kept for the composite entry points - GRU layer, HMR
 * regressor, blend - and as the Python convenience, which releases entries when the owning tensor dies or moves.) */
int gait_split_weight(const float* W_base, float* W_hilo, int64_t n, gait_stream_t stream);
int gait_linear_prepared(const float* A, int64_t lda, const float* W, int64_t ldw, const float* W_hilo, int64_t n_prepared,
                         const float* W_base, const float* bias, const float* Cin, int64_t ldcin, float* C, int64_t ldc,
                         int64_t M, int64_t N, int64_t K, gait_stream_t stream);

/* Debug hook: device buffer of 64*4 uint64 that receives per-k-block pipeline timestamps (stage free, data
 * landed, converted, MMAs issued) of CTA 0 of subsequent tensor-core GEMM launches; NULL disables. */
int gait_debug_linear_trace(unsigned long long* device_buffer);

/* ---- GRU: torch.nn.GRU as used by TemporalEncoder / gait_feat_encoder.py:51-57,88 --------- */
size_t gait_gru_workspace_bytes(int64_t S, int64_t T, int64_t H);
/* One layer, one direction.  x (S,T,I) with frame stride ldx; weights in torch layout
 * (W_ih (3H,I), W_hh (3H,H), gate order r,z,n); h0 (S,H) or NULL (zeros).
 * y  (S,T,.) frame stride ldy : raw GRU output h_t (written at y + (s*T+t)*ldy, H wide)
 * out (S,T,.) frame stride ldout, optional: h_t + resid[s,t,:]  (resid frame stride ldres; the
 *     TemporalEncoder residual); NULL to skip.
 * hn (S,H) optional final hidden state.  reverse!=0 runs t = T-1..0. */
int gait_gru_layer(const float* x, int64_t ldx, const float* W_ih, const float* W_hh, const float* b_ih,
                   const float* b_hh, const float* h0, float* y, int64_t ldy, const float* resid,
                   int64_t ldres, float* out, int64_t ldout, float* hn, int64_t S, int64_t T, int64_t I,
                   int64_t H, int reverse, void* workspace, size_t workspace_bytes, gait_stream_t stream);

/* Which recurrence implementation gait_gru_layer uses for (S,T,H) on the current device with 16-byte aligned operands and
 * H-wide strides: 2 = weight-stationary kernel (1-2 sequences), 1 = persistent cluster kernel (64-sequence launches),
 * 0 = one GEMM + gate kernel per time step.  Needs a device (queries occupancy and the kernels' register counts). */
int gait_gru_plan(int64_t S, int64_t T, int64_t H);

/* Debug hook: device buffer of 1280 uint64 receiving clock64 stamps of CTA 0 of the persistent recurrent kernel
 * (per step: [step*8+i]; per k-block of step 2: [256+kb*8+i]); NULL disables. */
int gait_debug_gru_trace(unsigned long long* device_buffer);

/* F.relu between the GRU and the optional output Linear of TemporalEncoder. y may alias x. */
int gait_relu(const float* x, float* y, int64_t n, gait_stream_t stream);

/* ---- HMR iterative regressor: spin.py:244-265 -------------------------------------------- */
/* state = [pose6d(144) | betas(10) | cam(3)] = 157 floats, stored with row stride 160.
 * W1x (Dh,Din) = fc1.weight[:, :Din];  W1s (Dh,160) = fc1.weight[:, Din:] zero padded;
 * W2 (Dh,Dh); Wd (157,Dh) = [decpose;decshape;deccam].weight; bd (157).
 * init (1,160) broadcast when init_rows==1, else (F,160).  state_out (F,160).  With a shared init the constant term
 * init . W1s^T is computed once per call and folded into the bias of the x-part GEMM (iteration 0 then needs no state GEMM). */
size_t gait_hmr_workspace_bytes(int64_t F, int64_t Dh);
int gait_hmr_regressor(const float* x, int64_t ldx, const float* W1x, const float* W1s, const float* b1,
                       const float* W2, const float* b2, const float* Wd, const float* bd,
                       const float* init, int64_t init_rows, int n_iter, float* state_out, int64_t F,
                       int64_t Din, int64_t Dh, void* workspace, size_t workspace_bytes,
                       gait_stream_t stream);

/* The same loop with the layers folded (opt-in).  spin.py:244-265 has no non-linearity between fc1, fc2 and the decoders and
 * dropout is the identity in eval(), so for the shared mean-parameter init the n_iter iterations are ONE affine map of x:
 * state = x . Wf^T + bf, Wf (157,Din), bf (157), folded by the caller in FP64 (gaitb200.regressor.Regressor.fold).
 * state_out (F,160), padding columns zeroed.  Same results as gait_hmr_regressor up to FP32 rounding. */
size_t gait_hmr_folded_workspace_bytes(int64_t F);
int gait_hmr_regressor_folded(const float* x, int64_t ldx, const float* Wf, const float* bf, float* state_out,
                              int64_t F, int64_t Din, void* workspace, size_t workspace_bytes, gait_stream_t stream);

/* ---- SMPL: smplx==0.1.26 lbs.py via lib/models/smpl.py:108-130 ---------------------------- */
/* Shape-dependent rest joints + 24-joint kinematic chain (smplx lbs.batch_rigid_transform), one
 * warp per frame, parent transforms exchanged by warp shuffle level by level.
 * R (F,24,3,3); betas (F,10) frame stride ldb; J_template (24,3) = J_regressor.v_template;
 * J_shapedirs (24,3,10) = J_regressor.shapedirs; parents int32[24] (parents[0] = -1).
 * A (F,24,12) optional: rows of the 3x4 skinning transform with the rest joint removed;
 * J_posed (F,24,3); coef (F,GAIT_BLEND_LD) optional: [R[1:]-I (207) | betas (10) | 1 | 0..];
 * Aop optional (gait_smpl_lbs_aop_bytes(F) bytes): the same transforms as the tensor-core LBS
 * operand (TF32 hi/lo split, UMMA core-matrix layout, one blob per 8 frames).  A or Aop required. */
int gait_smpl_pose_chain(const float* R, const float* betas, int64_t ldb, const float* J_template,
                         const float* J_shapedirs, const int32_t* parents, float* A, float* J_posed,
                         float* coef, float* Aop, int64_t F, gait_stream_t stream);
/* The same kernel with the 6-D -> rotation conversion (geometry.py:395-410, eps) in front and theta packing (spin.py:288,
 * pare.py:79: [cam | axis-angle(72) | betas], geometry.py:68-97 route) behind it, so the head needs one launch where
 * gait_rot6d_to_rotmat + gait_smpl_pose_chain + gait_pack_theta need three.  x6: row f at x6 + f*ldx6 holds the 24
 * interleaved 6-vectors; R_out (F,24,3,3) receives the rotations; theta (F,85) optional (needs cam (F,3), stride ldcam). */
int gait_smpl_pose_chain_rot6d(const float* x6, int64_t ldx6, float eps, const float* betas, int64_t ldb, const float* cam,
                               int64_t ldcam, const float* J_template, const float* J_shapedirs, const int32_t* parents,
                               float* R_out, float* A, float* J_posed, float* coef, float* Aop, float* theta, int64_t F,
                               gait_stream_t stream);
/* Blend shapes (smplx lbs.blend_shapes + pose offsets): v_posed (F,3V) = coef (F,224) . basis_t^T,
 * basis_t (3V,224) = [posedirs^T | shapedirs | v_template | 0]; v_posed row stride ldv >= 3V. */
int gait_smpl_blend(const float* coef, const float* basis_t, float* v_posed, int64_t ldv, int64_t F, int64_t V3,
                    gait_stream_t stream);
/* Linear blend skinning (last two lines of smplx lbs): verts[f,v] = (sum_j W[v,j] A[f,j]) [v_posed;1].
 * v_posed (F,V,3) with frame stride ldv; A (F,24,12); lbs_weights (V,24); verts (F,V,3).  SIMT FP32. */
int gait_smpl_lbs(const float* v_posed, int64_t ldv, const float* A, const float* lbs_weights, float* verts,
                  int64_t F, int64_t V, gait_stream_t stream);
/* The same skinning with the W.A contraction on tcgen05 tensor cores (FP32-accurate split TF32),
 * TMA bulk copies of pre-arranged operands, 128 vertices x 8 frames per CTA:
 *   Wpack  = gait_smpl_lbs_pack(lbs_weights)  (gait_smpl_lbs_pack_bytes(V) bytes, once per model)
 *   Aop    = the pose-chain kernel's tensor operand output
 *   v_posed rows padded: ldv >= 384*ceil(V/128), ldv % 4 == 0 (rows are bulk-copied 1536 B at a time)
 *   verts (F,V,3) contiguous, 8-byte aligned, V even
 *   jx (V) optional: one joint-regressor row (the MPII thorax row of J_regressor_extra); its dot
 *   product with the skinned vertices is emitted as partial sums over 32-vertex slices,
 *   jx_partial (gait_smpl_lbs_jx_parts(V) = 4*ceil(V/128), F, 3), summed by gait_joints_assemble (extra_parts = that). */
int64_t gait_smpl_lbs_jx_parts(int64_t V);
size_t gait_smpl_lbs_pack_bytes(int64_t V);
int gait_smpl_lbs_pack(const float* lbs_weights, float* packed, int64_t V, gait_stream_t stream);
size_t gait_smpl_lbs_aop_bytes(int64_t F);
int gait_smpl_lbs_tc(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                     float* verts, float* jx_partial, int64_t F, int64_t V, gait_stream_t stream);
/* Joints-only variant (BASELINE config 5, "no mesh write-back"): the same skinning, but the mesh is never written to
 * HBM.  Only the n_lm landmark vertices lm_idx[] (int32 vertex ids; the VertexJointSelector landmarks the joint sets
 * use) go to lm_out (F, n_lm, 3), plus the fused regressor-row partials; gait_joints_assemble then takes lm_out as its
 * `verts` with V = n_lm and landmarks = 0..n_lm-1. */
int gait_smpl_lbs_tc_joints(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                            float* jx_partial, const int32_t* lm_idx, int n_lm, float* lm_out, int64_t F, int64_t V,
                            gait_stream_t stream);
/* General form: verts (F,V,3) and/or lm_out (F,n_lm,3) (either may be NULL, not both).  With both, the mesh goes to `verts`
 * - which may be a PEER address (gait_peer_open): the kernel's coalesced stores then are the final gather of a
 * sequence-sharded run - while the landmark vertices the joint sets read stay in local memory. */
int gait_smpl_lbs_tc_ex(const float* v_posed, int64_t ldv, const float* Aop, const float* Wpack, const float* jx,
                        float* verts, float* jx_partial, const int32_t* lm_idx, int n_lm, float* lm_out, int64_t F,
                        int64_t V, gait_stream_t stream);
/* Joints-only path without the mesh (BASELINE configs[4]).  Skinning is linear in v_posed for fixed transforms, so the one
 * regressor row the Kinect-25 set needs (thorax) is  sum_j A[f,j] [P_j coef[f]; s_j]  with P_j (3,224) = sum_v jx[v] W[v,j]
 * basis[3v..3v+2,:] and s_j = sum_v jx[v] W[v,j] prepared once per model; the landmark vertices are skinned individually.
 * u (F, ldu >= 3 n_lm + 72) = coef . [basis rows of the landmark vertices (3 n_lm) ; P (72)]^T comes from gait_linear;
 * A (F,24,12) from the pose chain; lm_weights (n_lm,24) = lbs_weights[landmarks]; s (24).
 * lm_out (F,n_lm,3) and thorax (F,3) feed gait_joints_assemble (verts = lm_out, V = n_lm, one extra part). */
int gait_smpl_reduced_joints(const float* A, const float* u, int64_t ldu, const float* lm_weights, const float* s,
                             float* lm_out, float* thorax, int64_t F, int n_lm, gait_stream_t stream);
/* vertices2joints (smplx lbs; lib/models/smpl.py:113, pare.py:70-76, spin.py:279-282):
 * out (F,Rj,3) = Jreg (Rj,V) . verts (F,V,3). */
int gait_joint_regress(const float* verts, const float* Jreg, float* out, int64_t F, int64_t V, int Rj,
                       gait_stream_t stream);
/* The same regression as ONE streaming pass over the mesh (lane = frame, all Rj rows per pass, 8-CTA clusters splitting the
 * vertex range, partial sums combined through distributed shared memory; jreg.cu).  The regressor is packed once
 * (gait_joint_regress_pack -> gait_joint_regress_pack_bytes(V,Rj) bytes, 16-byte aligned: [row block][4-vertex group][row][4],
 * zero padded) so that every pipeline stage stages its weights with one TMA bulk copy.  verts (F,V,3) contiguous, V even. */
size_t gait_joint_regress_pack_bytes(int64_t V, int Rj);
int gait_joint_regress_pack(const float* Jreg, float* packed, int64_t V, int Rj, gait_stream_t stream);
int gait_joint_regress_packed(const float* verts, const float* packed, float* out, int64_t F, int64_t V, int Rj,
                              gait_stream_t stream);
/* Joint assembly (smplx VertexJointSelector + smpl.py:114-121) with optional projection
 * (smpl.py:176-186 / geometry.py:412-425) and Kinect-25 gather (kp_utils.py:26-36).
 * Virtual joint v: v<24 -> J_posed; 24<=v<24+n_landmarks -> verts[landmark[v-24]];
 * else -> extra[v-24-n_landmarks], extra (extra_parts, F, n_extra, 3) summed over its leading axis
 * (parts extra_part_stride floats apart; 1 part for a complete regression).  joint_map int32[J] picks virtual joints.
 * joints (F,J,3); kp2d (F,J,2) optional (needs cam (F,3), frame stride ldcam):
 *   t = [tx, ty, 2 f/(res s + 1e-9)], kp2d = f (X+t).xy/(X+t).z / kp2d_divisor;
 * gather int32[n_gather] + gathered (F,n_gather,3) optional (entry -1 writes zeros). */
int gait_joints_assemble(const float* J_posed, const float* verts, int64_t V, const int32_t* landmarks,
                         int n_landmarks, const float* extra, int n_extra, int extra_parts, int64_t extra_part_stride,
                         const int32_t* joint_map, int J,
                         float* joints, const float* cam, int64_t ldcam, float focal_length, float img_res,
                         float kp2d_divisor, float* kp2d, const int32_t* gather, int n_gather,
                         float* gathered, int64_t F, gait_stream_t stream);
/* convert_kps (kp_utils.py:26-36) as a device gather: dst (F,Jd,3)[f,k] = src (F,Js,3)[f,idx[k]],
 * zeros where idx[k] < 0 (a destination joint the source layout lacks). */
int gait_gather_joints(const float* src, int Js, const int32_t* idx, int Jd, float* dst, int64_t F,
                       gait_stream_t stream);
/* theta packing (spin.py:288, pare.py:79): theta (F,85) = [cam(3) | axis-angle(72) | betas(10)],
 * axis-angle from R (F,24,3,3) by the geometry.py:68-97 route. */
int gait_pack_theta(const float* R, const float* cam, int64_t ldcam, const float* betas, int64_t ldb,
                    float* theta, int64_t F, gait_stream_t stream);

/* ---- peer memory: the final gather of a sequence-sharded run ---------------- */
/* The one exchange on the path is the gather of meshes / Kinect-25 joints onto one rank; it replaces the per-process
 * device-to-host boundary batch_generation.py:316-323 / demo.py:183-188.  The root allocates the gathered buffer
 * (gait_peer_alloc: the one place this library allocates, because a CUDA IPC handle maps a whole allocation), exports a
 * 64-byte handle that the host side ships to the other processes (any transport), and they map it (gait_peer_open).
 * Ranks then write their blocks either with gait_smpl_lbs_tc_ex(verts = mapped address) or with gait_peer_copy
 * (asynchronous device-to-device copy on `stream`; dst/src may be local or mapped).  Completion is stream-ordered on the
 * writer; the host side signals the root (a collective or event) before it reads. */
#define GAIT_PEER_HANDLE_BYTES 64
int gait_peer_alloc(void** ptr, size_t bytes);
int gait_peer_free(void* ptr);
int gait_peer_export(const void* ptr, unsigned char handle[GAIT_PEER_HANDLE_BYTES]);
int gait_peer_open(const unsigned char handle[GAIT_PEER_HANDLE_BYTES], void** ptr);
int gait_peer_close(void* ptr);
int gait_peer_copy(void* dst, const void* src, size_t bytes, gait_stream_t stream);

/* ---- post-processing next to the head ------------------------------ */
/* One-Euro filter (lib/utils/one_euro_filter.py:5-46) over T frames of C channels, as lib/utils/smooth_pose.py:51-56,
 * 84-88 drives it: x (T,C) sampled at t = 0,1,2,...; x_hat[0] = x[0], dx_0 = 0.  numpy float32 arithmetic, bit-exact. */
int gait_one_euro_filter(const float* x, float* x_hat, int64_t T, int64_t C, double min_cutoff, double beta, double d_cutoff,
                         gait_stream_t stream);
/* convert_crop_cam_to_orig_img (lib/utils/demo_utils.py:176-193): cam (N,3) float32 [s,tx,ty] of the crop, bbox (N, ldb>=3)
 * [cx, cy, h, ..] float32 or float64 -> out (N,4) [sx, sy, tx, ty] in bbox's dtype (numpy promotion). */
int gait_crop_cam_to_orig_img(const float* cam, const void* bbox, int bbox_is_f64, int64_t ldb, double img_width,
                              double img_height, void* out, int64_t N, gait_stream_t stream);
/* convert_crop_coords_to_orig_img (lib/utils/demo_utils.py:196-209): keypoints (N,J,D>=2) float32 in [-1,1] crop units ->
 * out (N,J,D) float32 original-image pixels (may alias keypoints). */
int gait_crop_coords_to_orig_img(const void* bbox, int bbox_is_f64, int64_t ldb, const float* keypoints, float* out, int64_t N,
                                 int J, int D, double crop_size, gait_stream_t stream);

/* ---- heads next to the path ---------------------------------------- */
/* KeypointAttention (lib/models/layers/keypoint_attention.py:34-55, act='softmax'): heat (B,J,HW) is scaled, soft-maxed over
 * its HW pixels and used to pool feat (B,C,HW):  out[b*sob + c*soc + j*soj] = sum_hw softmax(heat[b,j])[hw] feat[b,c,hw]. */
int gait_keypoint_attention(const float* feat, const float* heat, float scale, float* out, int64_t B, int C, int J, int HW,
                            int64_t sob, int64_t soc, int64_t soj, gait_stream_t stream);
/* LocallyConnected2d (lib/models/layers/locallyconnected2d.py:39-49) with kernel_size 1 and output_size [J,1]:
 * out[n,o,j] = sum_c x[n,c,j] W[o,c,j] + bias[o,j]; every operand is addressed through element strides (s*), so a
 * broadcast input (sxj = 0, gait_feat_encoder.py:88), a weight shared by all joints (swj = 0: a 1x1 convolution) and a
 * transposed output need no copies.  Optional: out2 = out + resid (same layout as out). */
int gait_locally_connected(const float* x, int64_t sxn, int64_t sxc, int64_t sxj, const float* W, int64_t swo, int64_t swc,
                           int64_t swj, const float* bias, int64_t sbo, int64_t sbj, float* out, int64_t son, int64_t soo,
                           int64_t soj, const float* resid, float* out2, int64_t N, int C, int O, int J, gait_stream_t stream);
/* kind 0: LeakyReLU(slope), kind 1: Tanh (gait_feat_encoder.py:58-78). y may alias x. */
int gait_activation(const float* x, float* y, int64_t n, int kind, float slope, gait_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GAITB200_H */
