/*
 * pronerf_b200 -- C ABI of the B200-native (sm_100a) ProNeRF per-ray render hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every pointer named
 * "device" is a CUDA device pointer (e.g. torch.Tensor.data_ptr()); every launch is asynchronous
 * on the given stream (pass torch.cuda.current_stream().cuda_stream, or 0 for the legacy stream).
 * All tensors are dense row-major fp32 unless stated.  Functions return PN_OK (0) or a negative
 * PN_E* code; pn_last_error() gives the message of the calling thread's last failure.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the upstream
 * repository KAIST-VICLab/pronerf; "trt.py" = run_S_eS_eN_alter_trt.py, "helpers.py" =
 * run_nerf_helpers.py, "iw.py" = inverse_warp.py, "engines" = trt_infer_v2.py -- the reference's
 * own backend plug-in seam, whose MMEngine/RefineEngine/NeRFEngine objects own persistent device
 * buffers exactly like pn_ctx_t does).
 */
#ifndef PRONERF_B200_H
#define PRONERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PN_VERSION 100            /* 0.1.0 */

#define PN_OK            0
#define PN_EINVAL       -1       /* bad argument (shape, null pointer, unsupported size) */
#define PN_ECUDA        -2       /* a CUDA runtime call or launch failed */
#define PN_ENODEVICE    -3       /* no sm_100 device */
#define PN_ESTATE       -4       /* context not ready (weights not loaded, ...) */
#define PN_ENOMEM       -5

typedef void* pn_stream_t;        /* cudaStream_t */
typedef struct pn_ctx pn_ctx_t;   /* owns packed weights + scratch, like the reference's engine objects */

/* which network a weight set belongs to */
#define PN_NET_SAMPLER  0         /* MinMaxRaySamplerTRT_Net     helpers.py:1473-1507 */
#define PN_NET_REFINE   1         /* MinMaxRayEpiSamplerTRT_Net  helpers.py:1509-1540 */
#define PN_NET_NERF     2         /* DoNeRFTRT                   helpers.py:1186-1343 */

/* arithmetic tier of the three MLPs */
#define PN_PREC_FP32    0         /* fp32 SIMT FMA: the <=1e-3 max-abs parity tier            */
#define PN_PREC_F16     1         /* tensor-core tier: IEEE fp16 operands (max 65504), fp32 accumulate in TMEM on tcgen05    */
#define PN_PREC_BF16    PN_PREC_F16   /* deprecated alias (round-1 name); the operands are fp16, NOT bfloat16                */

int         pn_version(void);
const char* pn_last_error(void);
/* PN_OK iff `device` exists and is compute capability 10.x (the only target; no fallback). */
int         pn_device_check(int device);
/* 1 when the PN_PREC_BF16 (tcgen05) tier of the MLPs is compiled into this build, else 0. */
int         pn_has_bf16_tier(void);
/* Debug aid (not part of the render path): when set to a device buffer of 208 int64, CTA 0 of every subsequent
 * bf16 MLP launch records clock64() stamps of its second tile's pipeline events; NULL switches it off. */
int         pn_debug_tc_timeline(void* dev_buf_208_i64);
/* Measurement aid (not part of the render path): when set to a device buffer of 12 int64, CTA 0 of every subsequent tensor-core
 * MLP launch writes {clock64(), %globaltimer [ns]} after its setup and before its teardown into slots [4 net .. 4 net + 3]
 * (net: 0 sampler, 1 refine, 2 NeRF): SM cycles / wall time = the clock the launch really ran at (the board's power limit
 * holds these kernels at 1.45-1.8 GHz while nvidia-smi's averaged reading still says 1965 MHz).  NULL switches it off. */
int         pn_debug_tc_clock(void* dev_buf_12_i64);

/* ---- context: packed weights + scratch ------------------------------------------------------- */
int  pn_ctx_create(int device, pn_ctx_t** out);
void pn_ctx_destroy(pn_ctx_t* ctx);

/* Load one network from nn.Linear-layout device tensors: W[l] is [out_dims[l], in_dims[l]] row-major,
 * b[l] is [out_dims[l]].  Layer order = fc_backbone.0..5, fc_output (sampler / refine) or
 * layers.0..7 (DoNeRFTRT); state_dict keys per trt.py:478-481.  Hidden width must be 256.
 * Packs both the fp32 (k-major) and the 16-bit (fp16, UMMA canonical, 128B-swizzled) images of the weights. */
int pn_ctx_load_net(pn_ctx_t* ctx, int net, int n_layers, const int* in_dims, const int* out_dims,
                    const float* const* W_device, const float* const* b_device, pn_stream_t stream);

/* The classic NeRF (run_nerf_helpers.py:792-847: D=8, W=256, skips=[4], use_viewdirs, 63 + 27 encoded inputs) as the shading
 * network of this context -- the topology stage-2 checkpoints hold under 'network_fine_state_dict' (refine2.py:360-362, 890),
 * which the reference's own infer script cannot load into its DoNeRFTRT (trt.py:434-435, 481).  12 nn.Linear tensors in
 * checkpoint order: pts_linears.0..7, alpha_linear, feature_linear, views_linears.0, rgb_linear.  Afterwards pn_nerf_forward,
 * pn_run_network and pn_render_rays use it; it runs in PN_PREC_FP32 only (other precisions fail with PN_ESTATE).  Loading
 * either kind of shading network replaces the other. */
int pn_ctx_load_nerf_classic(pn_ctx_t* ctx, const int* in_dims, const int* out_dims, const float* const* W_device,
                             const float* const* b_device, pn_stream_t stream);

/* Stage timing for the roofline report: when enabled, pn_render_rays brackets each of its PN_N_STAGES kernels
 * with CUDA events on the launch stream (a ring of PN_PROFILE_RING frames; ~1 us per event).
 * pn_ctx_profile_read synchronises on the last recorded event, writes up to max_frames rows of
 * PN_N_STAGES per-kernel durations in ms (oldest first), returns the number of rows and clears the ring.
 * Stage order: sampler MLP, sort+lift, refine-Pluecker, project+gather, refine MLP, interval refine,
 * encode+NeRF MLP (incl. its view-direction pre-pass), composite.  On the PN_PREC_BF16 tier stages 2-4 are ONE kernel
 * (pn_refine_input_f16): its time is reported under project+gather and the other two read ~0.
 * (The reference times whole render() calls only, trt.py:327-332.) */
#define PN_N_STAGES      8
#define PN_PROFILE_RING  256
int pn_ctx_profile(pn_ctx_t* ctx, int enable);
int pn_ctx_profile_read(pn_ctx_t* ctx, float* ms, int max_frames);

/* ---- per-stage entry points (each mirrors one reference call) --------------------------------- */

/* MinMaxRaySamplerTRT_Net.forward (helpers.py:1490-1507) / MMEngine.run (engines:214-229).
 * x [N, in_dim] -> out [N, 3S+3] with the heads already applied:
 * [0,S) sigmoid = depth_values, [S,2S) mm_density_add, [2S,3S) mm_density_mul, [3S,3S+3) sigmoid = mm_rgb. */
int pn_sampler_forward(pn_ctx_t* ctx, const float* x, int64_t N, int S, float* out, int precision, pn_stream_t stream);

/* The same with the sampler input generated inside the kernel from the NDC ray batch (trt.py:274-278 fused: the 6P-wide
 * mm_input never exists): rays [N, ray_stride] (o_ndc, d_ndc, ...), P points per ray. */
int pn_sampler_forward_rays(pn_ctx_t* ctx, const float* rays, int ray_stride, int64_t N, int S, int P, float* out, int precision,
                            pn_stream_t stream);

/* MinMaxRayEpiSamplerTRT_Net.forward (helpers.py:1526-1540) / RefineEngine.run (engines:303-311).
 * x [N, 6S+3*NN*S] -> out [N, 4S+3]: [0,S) sigmoid = refine_depth, [S,4S) tanh = points_offset,
 * [4S,4S+3) sigmoid = refine_rgb. */
int pn_refine_forward(pn_ctx_t* ctx, const float* x, int64_t N, int S, float* out, int precision, pn_stream_t stream);

/* DoNeRFTRT.forward(input_pts [M,63], input_views [M,27]) -> raw [M,4]  (helpers.py:1331-1343) /
 * NeRFEngine.run (engines:385-394). */
int pn_nerf_forward(pn_ctx_t* ctx, const float* embedded, const float* embedded_dirs, int64_t M, float* raw,
                    int precision, pn_stream_t stream);

/* run_network (trt.py:195-208) with both positional encodings fused into the first / last layer:
 * pts [N,S,3], viewdirs [N,3] (row stride viewdir_stride floats: 3 for a dense tensor, 11 for
 * ray_batch[:, 8:11]) -> raw [N,S,4].  The encodings never touch HBM. */
int pn_run_network(pn_ctx_t* ctx, const float* pts, const float* viewdirs, int viewdir_stride, int64_t N, int S,
                   float* raw, int precision, pn_stream_t stream);

/* Embedder.embed via get_embedder(multires) (helpers.py:654-692): x [M,3] -> [M, 3+6L]. */
int pn_embed(const float* x, int64_t M, int L, float* out, pn_stream_t stream);

/* Pluecker.forward (helpers.py:629-632): o,d [M,3] -> [M,6] = [normalize(d), o x normalize(d)]. */
int pn_pluecker(const float* o, const float* d, int64_t M, float* out, pn_stream_t stream);

/* per-view sampler input (trt.py:274-278): rays [N,>=6] (o_ndc, d_ndc, ...) with row stride `ray_stride`
 * -> mm_input [N, 6P], P points linearly spaced on t in [0,1]. */
int pn_sampler_input(const float* rays, int ray_stride, int64_t N, int P, float* mm_input, pn_stream_t stream);

/* trt.py:631-637: scale by the per-ray (near, far) at rays[:,6:8], stable ascending sort of the S depths,
 * gather add/mul with the permutation, lift depth3d = 1/(1 - depth - 1e-5).
 * heads [N, >=3S] = sampler output (depth at [0,S), add at [S,2S), mul at [2S,3S)), row stride head_stride.
 * Outputs [N,S] each; perm is int32 (the reference's int64 permutation, narrowed). Any output may be NULL. */
int pn_sort_lift(const float* heads, int head_stride, const float* rays, int ray_stride, int64_t N, int S,
                 float* depth, float* add, float* mul, int32_t* perm, float* depth3d, pn_stream_t stream);

/* inverse_warp_rod1_rt2_coords_trt(img, depth, ro1, rd1, w2c, padding_mode='zeros') (iw.py:584-619).
 * img [B,C,H,W]; depth [B,N]; ro1, rd1 [*,4,N] with batch stride ro_bstride elements (0 for the
 * reference's expanded views); w2c [B,3,4] -> out [B,C,N].
 * Optional x0y0 [B,N,2] int32 receives floor(ix), floor(iy) (the integers that must be bit-exact;
 * clamped to +-2^30, non-finite -> -2^30). */
int pn_warp(const float* img, int B, int C, int H, int W, const float* depth, const float* ro1, const float* rd1,
            int64_t ro_bstride, const float* w2c, int64_t N, float* out, int32_t* x0y0, pn_stream_t stream);

/* Stage-2 TRAINING warp inverse_warp_rod1_rt2_coords(img, depth, ro1, rd1, c2w2, intrinsics, intrinsics_inv, scale=1,
 * padding_mode='zeros') (iw.py:515-581): like pn_warp but the source camera pose c2w2 [B,3,4] is inverted in the kernel, the
 * projection divides by |z| + 1e-8 and out-of-range normalised coordinates are pushed out of the image.  ro1, rd1 [*,3,N]
 * (batch stride ro_bstride, 0 for repeated views); intrinsics [B,3,3] (intrinsics_inv is unused by the reference too). */
int pn_warp_train(const float* img, int B, int C, int H, int W, const float* depth, const float* ro1, const float* rd1,
                  int64_t ro_bstride, const float* c2w2, const float* intrinsics, int64_t N, float* out, int32_t* x0y0,
                  pn_stream_t stream);

/* refine2.py:616-626: per-ray choice of NN source views (ref_nos [N,NN] int32) out of k_ref warped ones (warps [k_ref*S,3,N]),
 * warps that fell outside their source image replaced by the mean over the ray's valid views -> epi_features [N, 3*S*NN].
 * sample_major = 0: feature index (k*S+s)*3+ch (stage 2, refine2.py:626, and the infer path); 1: s*(NN*3)+k*3+ch (stage 1,
 * base.py:664-665). */
int pn_epi_features_train(const float* warps, const int32_t* ref_nos, int k_ref, int NN, int S, int64_t N, int sample_major,
                          float* epi, pn_stream_t stream);

/* Pack NN reference views [NN,H,W,3] (render_kwargs['images'][ref_nos], trt.py:286) into 16-byte RGBA fp32
 * texels [NN,H,W,4] so that one bilinear tap is one 128-bit load. */
int pn_pack_images(const float* images_hwc, int NN, int H, int W, float* texels, pn_stream_t stream);

/* trt.py:649-655 fused: lift + project every (neighbour k, sample s) of every ray and gather bilinearly.
 * texels from pn_pack_images; tex_index_host[k] (HOST ints, NULL = identity) names the texel image neighbour k
 * reads -- the per-view ref_nos ordering (trt.py:281-286) without moving pixels; project_mat [NN,3,4] (K * diag(1,-1,-1) * pose, trt.py:287-294);
 * ro_w, rd_w [N,3] world-space ray origin / direction with row stride ray_stride floats (3 for dense
 * tensors, 11 for or_ray_batch[:, 0:3] / [:, 3:6]); depth3d [N,S].
 * epi [N, epi_stride] gets feature (k*S+s)*3+ch at column epi_col0 + ... (epi_stride=3*NN*S, epi_col0=0 for
 * the reference's epi_features; epi_stride=6S+3*NN*S, epi_col0=6S writes straight into refine_input).
 * Optional x0y0 [NN*S, N, 2] int32 as in pn_warp. */
int pn_project_gather(const float* texels, const int* tex_index_host, int NN, int H, int W, const float* project_mat,
                      const float* ro_w, const float* rd_w, int ray_stride, const float* depth3d, int64_t N, int S,
                      float* epi, int epi_stride, int epi_col0, int32_t* x0y0, pn_stream_t stream);

/* trt.py:631-661 fused, for the PN_PREC_BF16 tier (S in {4, 8, 16}): pn_sort_lift + pn_refine_pluecker +
 * pn_project_gather in one kernel.  heads [N, >=3S] = sampler output; rays / or_rays [N, ray_stride] = NDC / world
 * ray batches (cols 0..7 / 0..5 used).  Outputs: depth, add, mul [N,S] (sorted, fp32) and refine_in_f16
 * [N, 6S+3*NN*S] = torch.cat([plucker_embed, epi_features]) (trt.py:661) rounded to IEEE fp16, 16-byte aligned,
 * which pn_refine_forward_f16 / the fused render path consume.  Optional x0y0 [NN*S, N, 2] int32 as in pn_warp. */
int pn_refine_input_f16(const float* heads, int head_stride, const float* rays, const float* or_rays, int ray_stride,
                        const float* texels, const int* tex_index_host, int NN, int H, int W, const float* project_mat,
                        int64_t N, int S, float* depth, float* add, float* mul, void* refine_in_f16, int32_t* x0y0,
                        pn_stream_t stream);

/* pn_refine_forward on the tensor-core tier with the fp16 input rows of pn_refine_input_f16:
 * x_f16 [N, 6S+3*NN*S] fp16 (dense, 16-byte aligned, row length a multiple of 8) -> out [N, 4S+3] fp32. */
int pn_refine_forward_f16(pn_ctx_t* ctx, const void* x_f16, int64_t N, int S, float* out, pn_stream_t stream);

/* trt.py:656-658: per-sample Pluecker features of (o + d*depth_s, d) into out[:, 0:6S] (row stride out_stride). */
int pn_refine_pluecker(const float* rays, int ray_stride, const float* depth, int64_t N, int S, float* out,
                       int out_stride, pn_stream_t stream);

/* trt.py:671-681: interval refinement + 1e-2 * offsets.  refine_out [N, >=4S] (refine_depth at [0,S),
 * offsets at [S,4S), row stride refine_stride) -> z [N,S], query [N,S,3]. */
int pn_interval_refine(const float* rays, int ray_stride, const float* depth, const float* refine_out,
                       int refine_stride, int64_t N, int S, float* z, float* query, pn_stream_t stream);

/* raw2outputs (trt.py:564-597): raw [N,S,4], z [N,S], rays_d = rays[:,3:6] (row stride ray_stride; pass a
 * [N,3] tensor with ray_d_col=0, ray_stride=3 for the stand-alone function), add/mul [N,S]
 * -> rgb [N,3], depth [N]; optional disp [N], acc [N], weights [N,S] (NULL to skip). */
int pn_composite(const float* raw, const float* z, const float* rays, int ray_stride, int ray_d_col,
                 const float* add, const float* mul, int64_t N, int S, float* rgb, float* depth, float* disp,
                 float* acc, float* weights, pn_stream_t stream);

/* Stage-1 flavour of raw2outputs (run_S_eS_eN_alter_base.py:501-548): raw is clamped to +-raw_clamp first (the reference uses
 * 10; <= 0 = no clamp) and add / mul may both be NULL (the NeRF-only step composites without the sampler's density heads). */
int pn_composite_stage1(const float* raw, const float* z, const float* rays, int ray_stride, int ray_d_col, const float* add,
                        const float* mul, float raw_clamp, int64_t N, int S, float* rgb, float* depth, float* disp,
                        float* acc, float* weights, pn_stream_t stream);

/* Stage-1 exploration sampling (base.py:689-707, 730), deterministic forward variant: depth [N,S] (sorted) ->
 * z [N, S*n_mult] = d_s + (m/n_mult) * |d_s - d_{s+1}| (d_S := far = rays[:,7]) and query [N, S*n_mult, 3] = o + dir * z.
 * (The reference draws n_mult in [1, 64/S], the direction and an extra |N(0, 0.2)| jitter at random per step.) */
int pn_explore_samples(const float* rays, int ray_stride, const float* depth, int64_t N, int S, int n_mult, float* z,
                       float* query, pn_stream_t stream);

/* The same step as the stage-1 TRAINING forward runs it (base.py:689-730, randomize=True, train_sampler=False), with the
 * reference's random draws as inputs so that it is reproducible: n_mult (random.randint(1, 64/S), base.py:690-691),
 * dir1_forward (random.random() > 0.5, base.py:697: spread towards the next sample / far, else towards the previous / near),
 * noise [N, S*n_mult] = abs(normal(0,1)/5) clamped at 0.99 (base.py:716-718; NULL = no jitter), dir2_forward (base.py:719).
 * z [N, S*n_mult] (replicas sorted like torch.sort, base.py:709; the jittered values are not re-sorted), query = o + dir * z. */
int pn_explore_samples_rand(const float* rays, int ray_stride, const float* depth, int64_t N, int S, int n_mult, int dir1_forward,
                            const float* noise, int dir2_forward, float* z, float* query, pn_stream_t stream);

/* ---- per-view prep (trt.py:245-278; helpers.py:2705-2714, 2776-2793) ---------------------------- */
/* get_rays + viewdir normalise + ndc_rays for one H x W view.  c2w [3,4] row-major HOST floats,
 * K: fx, fy, cx, cy as doubles.  rays [N,11] = (o_ndc, d_ndc, near, far, viewdir); or_rays [N,11] =
 * (o_w, d_w, or_near, or_far, viewdir).  row0/nrows select a horizontal band (tile sharding); N = nrows*W. */
int pn_raygen(int H, int W, double fx, double fy, double cx, double cy, const float* c2w_host, float near_,
              float far_, float or_near, float or_far, int row0, int nrows, float* rays, float* or_rays,
              pn_stream_t stream);

/* ---- the whole path --------------------------------------------------------------------------- */
typedef struct pn_frame {
  const float* rays;          /* device [N,11]  NDC ray batch (trt.py:269-271)                         */
  const float* or_rays;       /* device [N,11]  world-space twin (trt.py:250-254); cols 0..5 are used   */
  const float* mm_input;      /* device [N,6P] or NULL: NULL = generate from rays (same arithmetic)     */
  const float* texels;        /* device [n_img,H,W,4] from pn_pack_images                               */
  int tex_index[8];           /* neighbour k reads texel image tex_index[k] (ref_nos order, trt.py:281-286) */
  const float* project_mat;   /* device [NN,3,4]                                                        */
  int64_t N;
  int S, NN, P, H, W;
  int precision;              /* PN_PREC_*                                                              */
  float* rgb;                 /* device [N,3]  out                                                      */
  float* depth;               /* device [N]    out                                                      */
  /* Multi-view batches -- render_path's loop over poses (trt.py:236-332) as ONE pass: rays of view v occupy rows
   * [v*rays_per_view, (v+1)*rays_per_view), N = n_views*rays_per_view, project_mat is [n_views,NN,3,4] and
   * tex_index_views (HOST ints [n_views][NN], NULL = tex_index for every view) holds each view's ref_nos order.
   * n_views = 0 or 1: a single view, exactly the fields above.  At most PN_MAX_VIEWS views per batch. */
  int n_views;
  int64_t rays_per_view;
  const int* tex_index_views;
  /* Optional cudaEvent_t (NULL = none): the first kernel that reads `texels` waits for it on `stream`, so the caller can
   * upload + pack the reference views on another stream while the sampler MLP (which does not need them) runs. */
  void* texels_ready;
  /* Banded multi-view batches written straight into a frame set (tile sharding): 0 = dense output rows (row r
   * of the batch -> row r of rgb / depth).  > 0: rgb / depth point at (view 0, this band's first ray) of a frame set laid out
   * [n_views][out_view_stride rays]; ray r of view v lands at row v*out_view_stride + r.  With rgb / depth inside a peer-mapped
   * frame (pn_peer_open) the compositing kernel's stores ARE the tile gather. */
  int64_t out_view_stride;
  /* Optional cudaEvent_t (NULL = none) recorded on `stream` right after the LAST kernel of the pass that reads `texels`: a caller
   * that re-uploads the reference views for the next pass (the reference does so per view, trt.py:286) can start the upload as
   * soon as this event has fired instead of after the whole pass. */
  void* texels_done;
} pn_frame_t;
#define PN_MAX_VIEWS 16

/* render_rays (trt.py:599-696): sampler -> sort/lift -> project+gather -> refine -> interval refinement
 * -> encode + NeRF MLP -> composite, all on `stream`, scratch owned by ctx (grown on demand). */
int pn_render_rays(pn_ctx_t* ctx, const pn_frame_t* frame, pn_stream_t stream);

/* Same, HOST-buffer flavour (the end-to-end plug-in call): c2w_host [3,4] floats; uploads the pose,
 * generates rays on the device, renders rows [row0,row0+nrows) and copies rgb [nrows*W,3] / depth [nrows*W]
 * back into the (ideally pinned) host buffers; synchronises the stream before returning. */
int pn_render_view_host(pn_ctx_t* ctx, int H, int W, double fx, double fy, double cx, double cy,
                        const float* c2w_host, const float* texels, const int* tex_index_host,
                        const float* project_mat_host, int NN, int S, int P, int precision, int row0, int nrows,
                        float* rgb_host, float* depth_host, pn_stream_t stream);

/* render_path (trt.py:223-363) for n_views poses in one pass, HOST buffers: c2w_host [n_views,3,4], tex_index_host
 * [n_views,NN] (NULL = identity), project_mat_host [n_views,NN,3,4]; full frames; rgb_host [n_views*H*W,3] and
 * depth_host [n_views*H*W] (ideally pinned).  Uploads poses + matrices, generates all rays on the device, runs the
 * pn_render_rays pass over the n_views*H*W rays, downloads the frames and synchronises the stream.  With PN_PREC_BF16
 * the pass runs as two chunks (the first a whole number of waves of the persistent MLP kernels) and the first chunk's
 * frames are downloaded on a second stream while the second chunk renders; results are bit-identical to one pass.
 * texels_ready_event: optional cudaEvent_t as in pn_frame_t.texels_ready (NULL = texels are ready on `stream`). */
int pn_render_views_host(pn_ctx_t* ctx, int H, int W, double fx, double fy, double cx, double cy, int n_views,
                         const float* c2w_host, const float* texels, const int* tex_index_host,
                         const float* project_mat_host, int NN, int S, int P, int precision, float* rgb_host,
                         float* depth_host, void* texels_ready_event, pn_stream_t stream);

/* Pipelined flavour for a serving loop: the same pass for rows [row0, row0+nrows) of every view, returning as soon as everything
 * is enqueued.  Host frames are laid out [n_views][host_view_stride rays] (0 = dense, nrows*W per view) and rgb_host / depth_host
 * point at (view 0, the band's first ray) -- a rank of a sharded frame downloads its band straight into the shared host frame.
 * Results go home on the context's download stream; *ticket identifies the call for pn_wait.  Two calls may be in flight per
 * context (device frames are double-buffered; a third call's compositing waits on the device for the download two calls back);
 * the caller keeps rgb_host / depth_host untouched until pn_wait(ticket) returns.  c2w_host / project_mat_host / tex_index_host
 * may be reused as soon as the call returns.  texels_ready_event / texels_done_event: as pn_frame_t.texels_ready / .texels_done. */
int pn_render_views_host_async(pn_ctx_t* ctx, int H, int W, double fx, double fy, double cx, double cy, int n_views,
                               const float* c2w_host, const float* texels, const int* tex_index_host,
                               const float* project_mat_host, int NN, int S, int P, int precision, int row0, int nrows,
                               float* rgb_host, float* depth_host, int64_t host_view_stride, void* texels_ready_event,
                               void* texels_done_event, pn_stream_t stream, int64_t* ticket);
/* Block until the frames of `ticket` (and of every earlier call) are in host memory. */
int pn_wait(pn_ctx_t* ctx, int64_t ticket);

/* ---- peer frame buffers: the final tile gather of a sharded frame as direct stores over NVLink -------------------------
 * One process per GPU.  The destination rank calls pn_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; handle64 = 64 bytes to
 * ship to the other ranks by any means), every other rank pn_peer_open's the handle (cudaIpcOpenMemHandle with lazy peer
 * access) and passes `mapped pointer + its band's offset` as pn_frame_t.rgb / .depth: the compositing kernel then writes the
 * band straight into the destination GPU's frame while it renders, and the only thing left is a barrier.  (The reference has
 * no distributed code)  pn_peer_close unmaps, pn_peer_free releases the owner's allocation. */
int pn_peer_alloc(int device, size_t bytes, void** dev_ptr, unsigned char* handle64);
int pn_peer_open(int device, const unsigned char* handle64, void** dev_ptr);
int pn_peer_close(void* dev_ptr);
int pn_peer_free(void* dev_ptr);
/* Completion flags of the peer-store gather, in the same peer-mapped allocation: after its band, rank r enqueues
 * pn_peer_signal(&flags[r], step) (a system-scope release store over NVLink, ordered after the compositing kernel's stores by
 * the stream); the destination enqueues pn_peer_wait(flags, n, step, ...), a one-warp kernel that spins (acquire loads) until
 * every flag is >= step -- the frame is then complete in its memory without any host round trip or collective.  A watchdog
 * (timeout_ms) writes 1 + the late rank's index to *status_dev and lets the stream go on instead of hanging the GPU.
 * step <= 0: the step number is kept on the device instead -- *counter_dev (local memory of the calling rank, zero-initialised;
 * one counter for the signals, another for the waits) is advanced by the kernel itself, so that a CUDA graph holding the whole
 * sharded step (render + signal + wait) can be captured once and replayed. */
int pn_peer_signal(int* flag_dev, int step, int* counter_dev, pn_stream_t stream);
int pn_peer_wait(const int* flags_dev, int n_flags, int step, int timeout_ms, int* status_dev, int* counter_dev, pn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PRONERF_B200_H */
