/* priorcorr.h — C ABI of libpriorcorr.so, the B200 (sm_100a) implementation of PriOr-RAFT's
 * correlation hot path.
 *
 * The reference (longliangLiu/PriOr-Flow) has no native/FFI layer: its hot path is eager PyTorch
 * and its one native seam, `alt_cuda_corr.forward` (PriOr-RAFT/core/corr.py:86), is not shipped.
 * The drop-in boundary is therefore the set of Python names `core/prior_raft.py` resolves
 * (SURVEY.md §8b); this header is the C ABI *underneath* those names.  Every entry point cites the
 * reference function it replaces (paths relative to PriOr-RAFT/).  INTEGRATION.md shows the
 * ctypes binding and the reference-side patch a maintainer would add.
 *
 * Conventions
 *   - plain C: raw device pointers + int shapes, no torch / CUDA types in any signature
 *     (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream);
 *   - all tensors are dense fp32, row-major, in the layouts written next to each field;
 *   - the library never allocates, frees or retains device memory: outputs and workspaces are
 *     allocated by the caller (PyTorch's caching allocator in the host layer);
 *   - every call is asynchronous on `stream`, re-entrant, and honours the current device;
 *   - return 0 on success; non-zero => pf_last_error() (thread-local) describes the failure.
 *     No abort(), no exceptions across the ABI, no implicit device synchronisation.
 */
#ifndef PRIORCORR_H_
#define PRIORCORR_H_

#ifdef __cplusplus
extern "C" {
#endif

#define PF_ABI_VERSION 5   /* 5: pf_lookup_onthefly_tc, pf_onthefly_absmax, pf_onthefly_split */
#define PF_MAX_LEVELS 4

/* `tensor / python_scalar`: IEEE division on CPU, multiply by fp32 reciprocal in ATen's CUDA
 * kernels.  Coordinates are bit-exact against the chosen flavour of the reference. */
enum pf_div_mode { PF_DIV_IEEE = 0, PF_DIV_ATEN_CUDA = 1 };

/* Arithmetic of the all-pairs contraction (core/prior_raft.py:73, fp32 cuBLAS in the reference). */
enum pf_volume_mode {
  PF_VOL_FP32_3XF16 = 0, /* tcgen05 kind::f16 on an fp16 hi/lo split of both operands, 3 MMAs   */
                         /* (hi*hi + hi*lo + lo*hi): fp32-equivalent, the default               */
  PF_VOL_F16 = 1,        /* tcgen05 kind::f16, hi*hi only: TF32-class accuracy, separately      */
                         /* toleranced fast mode                                                */
  PF_VOL_FP32_SIMT = 2   /* CUDA-core FFMA tiles: exact fp32 products, the on-device checker    */
};

int pf_abi_version(void);
const char *pf_last_error(void);
/* Compile-time facts of the build: "sm_100a;tcgen05;tma;abi=4". */
const char *pf_build_info(void);

/* ------------------------------------------------------------------------------------------------
 * (a) all-pairs correlation volume + average-pool pyramid, fused.
 * Replaces PriOr_RAFT.corr (core/prior_raft.py:69-75; twin CorrBlock.corr core/corr.py:53-61) and
 * DCCL.build_pyramid (core/corr.py:99-111; twin CorrBlock.__init__ core/corr.py:22-28).
 *   level0[b, n, m] = sum_c fmap1[b,c,n] * fmap2[b,c,m] / sqrt(C)      n = y1*w+x1, m = y2*w+x2
 *   level(l+1) = avg_pool2d(level l, 2, stride 2) over the (y2,x2) axes
 * The tcgen05 modes need w % 32 == 0, h % 8 == 0, (h*w) % 128 == 0 and C % 32 == 0.
 */
typedef struct pf_volume_args {
  int batch, channels, h, w;     /* fmaps are [B, C, h, w]                                         */
  int num_levels;                /* 1..PF_MAX_LEVELS                                               */
  int mode;                      /* enum pf_volume_mode                                            */
  const float *fmap1;            /* [B, C, h, w]                                                   */
  const float *fmap2;            /* [B, C, h, w]                                                   */
  float *level[PF_MAX_LEVELS];   /* out: level l is [B*h*w, h>>l, w>>l]                            */
  void *workspace;               /* tcgen05 modes: pf_volume_workspace_bytes() bytes, 1 KiB aligned */
  long long workspace_bytes;
} pf_volume_args;
long long pf_volume_workspace_bytes(int batch, int channels, int h, int w, int mode);
int pf_volume_build(const pf_volume_args *args, void *stream);

/* One 2x2 average-pool level on its own (core/corr.py:108): in [planes, H, W] -> out [planes, H/2, W/2]. */
int pf_avg_pool2x2(const float *in, float *out, long long planes, int H, int W, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (b) dual-cost pyramid lookup.
 * Replaces DCCL.__call__ (core/corr.py:113-144) — own-view 9x9 window lookup, the window mapped
 * through the level-0 rotation grid into the other view's volume, and the img_rotate of that map —
 * and, with other[] == NULL and cyclic == 0, CorrBlock.__call__ (core/corr.py:30-51).
 * Output channel = level*(2r+1)^2 + a*(2r+1) + b sampling (x + a - r, y + b - r)  (x-major window).
 */
typedef struct pf_lookup_args {
  int batch, h, w;                     /* query grid: coords are [B, 2, h, w] (x, y)              */
  int h2, w2;                          /* level-0 plane size of the pyramids                      */
  int radius, num_levels;
  int cyclic;                          /* 1: cycle_bilinear_sampler (x % W); 0: bilinear_sampler  */
  int div_mode;                        /* enum pf_div_mode                                        */
  const float *coords;                 /* [B, 2, h, w]                                            */
  const float *own[PF_MAX_LEVELS];     /* [B*h*w, h2>>l, w2>>l]                                   */
  const float *other[PF_MAX_LEVELS];   /* same shapes; all NULL => single-view lookup             */
  const float *grid_w2c;               /* [B, 2, h, w] `sample_grid_*_W2C_8x`                     */
  const float *grid_c2w;               /* [B, 2, h, w] `sample_grid_*_8x`                         */
  long long grid_batch_stride;         /* elements between batches of the grids (0 = shared)      */
  float *out_own;                      /* [B, L*(2r+1)^2, h, w]                                   */
  float *out_other;                    /* [B, L*(2r+1)^2, h, w]                                   */
  float *scratch;                      /* [B, L*(2r+1)^2, h, w] pre-rotation map (caller-owned)   */
  float *dbg_own_xy;                   /* optional [B*h*w, L, (2r+1)^2, 2] unnormalised (ix, iy)  */
  float *dbg_other_xy;                 /* optional, same shape, orthogonal branch                 */
  int out_channels_last;               /* 1: out_own / out_other are [B, h, w, L*(2r+1)^2]        */
  int fuse_sum;                        /* 1: out_own = own + other (core/prior_raft.py:187-188),  */
                                       /*    out_other is not written (may be NULL)               */
  float *scratch_own;                  /* optional [B, h, w, L*(2r+1)^2] (caller-owned).  With    */
                                       /* fuse_sum and NCHW output the own view is staged here    */
                                       /* channels-last and added in the rotate kernel's single   */
                                       /* write pass instead of a read-modify-write of out_own    */
  int no_rotate;                       /* 1 (needs out_channels_last): stop after the gather — out_own holds the own view and   */
                                       /* `scratch` the other view BEFORE img_rotate, both [B, h, w, L*(2r+1)^2]; the consumer  */
                                       /* is pf_dccl_conv, which rotates, sums and convolves in one kernel                      */
} pf_lookup_args;
int pf_lookup_dual(const pf_lookup_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (f1) img_rotate + `own + other` + the motion encoder's first layer, fused: replaces the tail of DCCL.__call__
 * (core/corr.py:137-138), `corr_A + corr_B_A` (core/prior_raft.py:187-188) and `F.relu(self.convc1_A(corr))` /
 * `F.relu(self.convc1(corr))` — Conv2d(324, 256, 1) (core/update.py:168,184 and :85,92) — with one tcgen05 kernel:
 *   out[b, o, n] = relu(bias[o] + sum_c W[o, c] * (own[b, n, c] + img_rotate(raw)[b, n, c]))
 * Inputs are what pf_lookup_dual leaves behind with no_rotate = 1.  The weights are pre-split once into fp16 planes with
 * pf_dccl_conv_prepare (pf_dccl_conv_weight_bytes() bytes, 1 KiB aligned; redo it when the weights change).
 * split = 1: fp16 hi/lo three-product scheme, fp32-class accuracy (stated 1e-5 of max|ref|); split = 0: single fp16
 * product, TF32-class accuracy (stated 2e-3) — what cuDNN runs for this layer under torch.backends.cudnn.allow_tf32. */
typedef struct pf_dccl_conv_args {
  int batch, h, w;                     /* query grid                                               */
  int in_channels, out_channels;       /* 324, 256                                                 */
  int div_mode;                        /* enum pf_div_mode (img_rotate's sampler)                  */
  int split;                           /* 1: fp32-class (3 products); 0: TF32-class (1 product)    */
  int out_channels_last;               /* 1: out is [B, h, w, 256]; 0: [B, 256, h, w]              */
  int after_lookup;                    /* 1: launched right behind pf_lookup_dual on the same stream (programmatic dependent launch) */
  const float *raw;                    /* [B, h, w, 324] other view before img_rotate (`scratch`)  */
  const float *own_cl;                 /* [B, h, w, 324] own view                                  */
  const float *grid_c2w;               /* [B, 2, h, w] `sample_grid_*_8x`                          */
  long long grid_batch_stride;
  const void *prepared_weight;         /* pf_dccl_conv_prepare output                              */
  const float *bias;                   /* [256]                                                    */
  float *out;
} pf_dccl_conv_args;
long long pf_dccl_conv_weight_bytes(void);
int pf_dccl_conv_prepare(const float *weight /*[256, 324]*/, int out_channels, int in_channels, void *prepared, void *stream);
int pf_dccl_conv(const pf_dccl_conv_args *args, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (c) on-the-fly lookup: same outputs as pf_lookup_dual but straight from the feature maps, no
 * volume in memory.  Takes the place of the unshipped `alt_cuda_corr.forward(fmap1, fmap2_l,
 * coords_l, r)` behind AlternateCorrBlock (core/corr.py:64-91) and keeps its operand convention:
 * CHANNELS-LAST feature maps (core/corr.py:82-83).  Level l correlates fmap1 with avg-pooled fmap2
 * (pooling is linear, so this equals the lookup into the pooled volume up to fp rounding).
 */
typedef struct pf_onthefly_args {
  int batch, channels, h, w;                 /* channels % 128 == 0, <= 512                       */
  int radius, num_levels;
  int cyclic, div_mode;
  const float *coords;                       /* [B, 2, h, w]                                      */
  const float *fmap1_own;                    /* [B, h, w, C] query features of this view          */
  const float *fmap2_own[PF_MAX_LEVELS];     /* [B, h>>l, w>>l, C] pooled target features         */
  const float *fmap1_other;                  /* other view (NULL => single view)                  */
  const float *fmap2_other[PF_MAX_LEVELS];
  const float *grid_w2c, *grid_c2w;
  long long grid_batch_stride;
  float *out_own, *out_other, *scratch;      /* as in pf_lookup_args                              */
} pf_onthefly_args;
int pf_lookup_onthefly(const pf_onthefly_args *args, void *stream);
/*
 * (c') The same lookup with the dot products on tcgen05 tensor cores (pf_onthefly_tc.cu).  Per tile of 8 x 16 queries the
 * bounding box of all taps is found first; every box up to 256 target columns wide gets a dense [128 queries x box]
 * contraction of pre-split fp16 hi/lo planes (three products, fp32 accumulation: the volume kernel's numerics) written to
 * "local planes" in a pool of 16 KiB segments (128 queries x 32 box pixels); boxes beyond the pool's end or wider than 256
 * columns take the CUDA-core path.  The pool is the O(N) scratch that replaces the O(N^2) volume: 2 segments per query and
 * view (32 KiB) hold the boxes of a smooth flow field several times over.  Windows across the ERP seam stay on the
 * tensor-core path: the target planes are stored twice side by side.  Restrictions: radius 4, cyclic, h % 8 == 0,
 * w % 16 == 0, channels % 128 == 0.  Prepare the planes once per pyramid:
 *   zero the view's two `amax` words; pf_onthefly_absmax(fmap1, .., amax); pf_onthefly_absmax(fmap2 level 0, .., amax + 1);
 *   pf_onthefly_split(fmap1, .., amax, hi, lo, 0); pf_onthefly_split(fmap2 level l, .., amax + 1, hi_l, lo_l, (w>>l) * C).
 */
/* ints of `worklist` for T = views * num_levels * batch * (h/8) * (w/16) query tiles and a pool of `segs` segments */
#define PF_OTF_WORK_INTS(T, segs) (16 + 10 * (long long)(T) + 2 + 2 * ((long long)(segs) / 8 + (long long)(T)))
#define PF_OTF_SEGMENT_BYTES 16384
typedef struct pf_onthefly_tc_args {
  pf_onthefly_args base;                      /* fp32 operands and outputs exactly as for pf_lookup_onthefly       */
  const void *f1_hi_own, *f1_lo_own;          /* [B, h, w, C] fp16 planes of fmap1_own                             */
  const void *f2_hi_own[PF_MAX_LEVELS], *f2_lo_own[PF_MAX_LEVELS];      /* [B, h>>l, 2 (w>>l), C] fp16: rows doubled */
  const void *f1_hi_other, *f1_lo_other;
  const void *f2_hi_other[PF_MAX_LEVELS], *f2_lo_other[PF_MAX_LEVELS];
  const void *amax_own, *amax_other;          /* uint32[2] per view: absmax bits of {fmap1, fmap2 level 0}          */
  float *pool;                                /* scratch: pool_segments * PF_OTF_SEGMENT_BYTES bytes, 128-B aligned */
  long long pool_segments;
  int *worklist;                              /* scratch: PF_OTF_WORK_INTS(T, pool_segments) ints, 16-B aligned     */
  float *tap_xy;                              /* scratch (dual lookups): L * B * h * w * 81 * 2 floats, 8-B aligned  */
  int no_rotate;                              /* 1: stop before img_rotate and leave BOTH views channels-last — out_own and
                                               * scratch as [B, h, w, L*81] — the inputs of pf_dccl_conv (own_cl, raw)   */
} pf_onthefly_tc_args;
int pf_lookup_onthefly_tc(const pf_onthefly_tc_args *args, void *stream);
int pf_onthefly_absmax(const float *x, long long count, void *amax_word, void *stream);
int pf_onthefly_split(const float *x, long long count, const void *amax_word, void *hi, void *lo, long long dup_row_elems, void *stream);
/* Pooled feature pyramid helper: fmap [planes, H, W] -> levels[1..L-1] (levels[0] is ignored). */
int pf_fmap_pyramid(const float *fmap, float *const *levels, int num_levels, long long planes, int H, int W,
                    void *stream);

/* ------------------------------------------------------------------------------------------------
 * (d) ERP <-> orthogonal-view geometry.
 */
/* generate_samplegrid (core/utils/projection_prim_ortho.py:432-443 with :10-20, :397-411, :77-89,
 * :247-261, :51-74, :413-429): out[b,0/1,n,m] = source pixel (m', n') of (m, n) under rotation R. */
int pf_samplegrid(float *out /*[B,2,H,W]*/, int batch, int H, int W, const float *R_host /*9 floats, row-major, HOST*/,
                  int div_mode, void *stream);

/* Bilinear remap, zeros padding, align_corners=True, optional x % W.  Replaces
 * cycle_bilinear_sampler (core/utils/utils.py:78-95), bilinear_sampler (:61-75) and img_rotate
 * (core/utils/projection_prim_ortho.py:507-514 via :119-135).  Coordinates are read as
 * x = coords[b*batch_stride + p*pixel_stride], y = coords[... + xy_stride], p = yo*Wo + xo, which covers
 * both [B,Ho,Wo,2] (pixel_stride 2, xy_stride 1) and [B,2,Ho,Wo] (pixel_stride 1, xy_stride Ho*Wo). */
typedef struct pf_remap_args {
  int batch, channels, H, W;       /* src [B, C, H, W]                                              */
  int Ho, Wo;                      /* out [B, C, Ho, Wo]                                            */
  int cyclic, div_mode;
  const float *src;
  const float *coords;
  long long coord_batch_stride, coord_pixel_stride, coord_xy_stride;
  float *out;
} pf_remap_args;
int pf_remap(const pf_remap_args *args, void *stream);

/* flo_rotate with both grids given (core/utils/projection_prim_ortho.py:531-546, flow2endpoint
 * :200-218, u_clip :234-244, cycle_grid_sample / adjust_sample_m core/utils/my_cycle_sample.py:6-97),
 * one fused launch.  flow, out: [B,2,H,W]; grids [B,2,H,W] with the given batch stride. */
int pf_flo_rotate(const float *flow, const float *grid_w2c, const float *grid_c2w, long long grid_batch_stride,
                  float *out, int batch, int H, int W, void *stream);

/* Feature warp + group-wise correlation (core/prior_raft.py:173-174 + :77-83), fused:
 * out[b,g,p] = mean_{c in group g} fmap1[b,c,p] * cycle_bilinear_sampler(fmap2, coords)[b,c,p]. */
int pf_warp_groupcorr(const float *fmap1, const float *fmap2, const float *coords /*[B,2,h,w]*/, float *out /*[B,G,h,w]*/,
                      int batch, int channels, int h, int w, int groups, int div_mode, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (e) backward kernels (training; coords and grids carry no gradient, core/prior_raft.py:171,176).
 */
/* d(lookup)/d(pyramids): scatter-add of grad_own / grad_other into zero-initialised (or running)
 * gradient pyramids with the forward's coordinates.  grad_other is first pushed through the
 * adjoint of img_rotate into `scratch`. */
typedef struct pf_lookup_bwd_args {
  pf_lookup_args fwd;                     /* same geometry/coords/grids as forward; own/other unused  */
  const float *grad_own, *grad_other;     /* [B, L*(2r+1)^2, h, w]                                     */
  float *dgrad_own[PF_MAX_LEVELS];        /* += ; [B*h*w, h2>>l, w2>>l]                                */
  float *dgrad_other[PF_MAX_LEVELS];
  int query_begin, query_count;           /* query_count > 0: scatter only queries [query_begin, +query_count) of every batch item;  */
                                          /* the gradient pyramids then hold query_count planes per batch item: [B*query_count, ..]  */
                                          /* (the chunked, volume-free backward of the on-the-fly lookup)                            */
  int scratch_ready;                      /* 1: fwd.scratch already holds the img_rotate adjoint of grad_other (an earlier chunk)    */
} pf_lookup_bwd_args;
int pf_lookup_dual_bwd(const pf_lookup_bwd_args *args, void *stream);

/* Adjoints of the volume contraction (autograd of core/prior_raft.py:73-75) on tcgen05, both in one launch:
 *   dfmap1[b,c,n] = sum_m dvolume[b,n,m] fmap2[b,c,m] / sqrt(C)      dfmap2[b,c,m] = sum_n dvolume[b,n,m] fmap1[b,c,n] / sqrt(C)
 * dvolume is the level-0 gradient after pf_pyramid_fold_bwd.  bf16 hi/lo split of both operands, three products, fp32
 * accumulation: relative error ~2^-16 (stated tolerance for gradients: 1e-4 of max|ref|).  Needs C == 256, h*w % 128 == 0. */
typedef struct pf_volume_bwd_args {
  int batch, channels, h, w;
  const float *fmap1, *fmap2;          /* [B, C, h, w]                                             */
  const float *dvolume;                /* [B, h*w, h*w]                                            */
  float *dfmap1, *dfmap2;              /* [B, C, h, w]; either may be NULL                         */
  void *workspace;                     /* pf_volume_bwd_workspace_bytes() bytes, 1 KiB aligned     */
  long long workspace_bytes;
  /* chunked use (the volume-free backward of the on-the-fly lookup): dvolume holds only the rows of queries
   * [query_begin, +query_count) — [B, query_count, h*w]; dfmap1 is written for those queries only, dfmap2 is accumulated.   */
  int query_begin, query_count;        /* query_count == 0: the whole volume; else multiples of 128 / 64                     */
  int accumulate_dfmap2;               /* 1: dfmap2 += (every chunk after the first)                                         */
  int planes_ready;                    /* 1: the workspace already holds the bf16 planes of fmap1 / fmap2 (an earlier chunk) */
} pf_volume_bwd_args;
long long pf_volume_bwd_workspace_bytes(int batch, int channels, int h, int w);
int pf_volume_bwd(const pf_volume_bwd_args *args, void *stream);

/* Adjoint of pf_remap w.r.t. src: dsrc += scatter(dout) (dsrc must be initialised by the caller). */
int pf_remap_bwd(const pf_remap_args *args, const float *dout, float *dsrc, void *stream);

/* Adjoint of the pyramid: folds the level gradients into level 0 in place,
 * g0[n,y,x] += g1[n,y/2,x/2]/4 + g2[n,y/4,x/4]/16 + g3[n,y/8,x/8]/64. */
int pf_pyramid_fold_bwd(float *const *glevel, int num_levels, long long planes, int H, int W, void *stream);

/* Adjoint of pf_warp_groupcorr w.r.t. fmap1 and fmap2 (dfmap2 accumulated with atomics, caller zeroes). */
int pf_warp_groupcorr_bwd(const float *fmap1, const float *fmap2, const float *coords, const float *dout,
                          float *dfmap1, float *dfmap2, int batch, int channels, int h, int w, int groups,
                          int div_mode, void *stream);

/* ------------------------------------------------------------------------------------------------
 * (f2) / (f4): the callers either side of the path that reuse its geometry.
 */
/* PriOr_RAFT.upsample_flow (core/prior_raft.py:58-67): flow [B,2,h,w], mask [B,576,h,w] (or channels-last [B,h,w,576];
 * channel = k*64 + di*8 + dj) -> out [B,2,8h,8w] = sum_k softmax_k(mask) * 8 * flow[3x3 neighbour k], zero padded. */
int pf_convex_upsample(const float *flow, const float *mask, float *out, int batch, int h, int w, int mask_channels_last,
                       void *stream);
/* Its adjoint: dflow [B,2,h,w] (zeroed here, accumulated with atomics: float sums in arbitrary order) and dmask in the mask's
 * layout, from grad_out [B,2,8h,8w]; the softmax is recomputed from `mask`. */
int pf_convex_upsample_bwd(const float *flow, const float *mask, const float *grad_out, float *dflow, float *dmask, int batch, int h, int w,
                           int mask_channels_last, void *stream);
/* One term of uniform_loss (train_flow.py:55-79): *acc += term_weight * sum(ok * lat[y] * |pred - gt|_1), with ok [B,H,W] the
 * 0/1 validity mask, lat [H] the normalised cos-latitude weights (core/utils/spherical.py:11-17); and its gradient
 * dpred = *upstream * term_weight * ok * lat[y] * sign(pred - gt). */
int pf_uniform_loss_fwd(const float *pred, const float *gt, const float *ok, const float *lat, float *acc, float term_weight,
                        int batch, int H, int W, void *stream);
int pf_uniform_loss_bwd(const float *pred, const float *gt, const float *ok, const float *lat, const float *upstream,
                        float term_weight, float *dpred, int batch, int H, int W, void *stream);
/* calculate_great_circle_distance(pred, gt, 'Haversine', R) (core/utils/spherical.py:20-53): [B,2,H,W] x2 -> [B,H,W]. */
int pf_great_circle(const float *pred, const float *gt, float *out, int batch, int H, int W, float radius, void *stream);

/* ------------------------------------------------------------------------------------------------
 * Measurement aids (bench.py): no reference counterpart.
 * pf_probe_gather issues exactly the loads of an own-view level-0 lookup (core/corr.py:128: per plane a 10x10 footprint at
 * pos_xy[2n], pos_xy[2n+1], ten coalesced 40-byte row segments) and nothing else; its time is the HBM floor of that access
 * pattern.  pf_probe_stream_read reads `count` floats once (read-only streaming ceiling).  `sink` is one float. */
int pf_probe_gather(const float *vol /*[planes,H,W]*/, long long planes, int H, int W, const int *pos_xy /*[planes,2]*/,
                    float *sink, void *stream);
int pf_probe_stream_read(const float *src, long long count, float *sink, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PRIORCORR_H_ */
