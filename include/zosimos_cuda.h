/* zosimos_cuda.h -- C ABI of the B200 execution backend for zosimos compositing programs.
 *
 * This is the drop-in boundary: the entry points a Rust `zosimos-cuda-sys` crate (or any FFI)
 * binds in place of the reference's wgpu encoder + executor.  Reference = 197g/zosimos; every
 * declaration cites the reference interface it replaces (paths relative to the reference root).
 * See INTEGRATION.md for the Rust-side binding.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns zos_status (0 = ok) and never
 *     aborts or unwinds (reference: Result<_, LaunchError|StartError|StepError>,
 *     lib/zosimos/src/program.rs:1997-2006, lib/zosimos/src/run.rs:370-408);
 *     zos_last_error(ctx) gives a human readable message for the last failure.
 *   - a zos_ctx owns one CUDA device + one stream; it is NOT thread safe (like `&mut Execution`,
 *     run.rs:1389); distinct contexts are independent (one per GPU / per host thread).
 *   - images live in pitch-linear device buffers; rows are padded to 256 bytes exactly like
 *     Descriptor::to_aligned (lib/zosimos/src/buffer.rs:121-134); use zos_aligned_row_stride.
 *   - numeric codes for transfer / sample parts / sample bits are the reference's own
 *     (lib/zosimos/src/shaders/stage.rs:74-119, lib/std/src/stage.frag:107-170).
 *   - there is NO CPU fallback: without a CUDA device every call fails with ZOS_ERR_CUDA.
 */
#ifndef ZOSIMOS_CUDA_H
#define ZOSIMOS_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZOS_ABI_VERSION 5

typedef int32_t zos_status;
enum {
  ZOS_OK = 0,
  ZOS_ERR_INVALID = 1,        /* bad argument / inconsistent descriptor (CommandErrorKind::BadDescriptor) */
  ZOS_ERR_UNSUPPORTED = 2,    /* valid but not implemented texel / op (CommandError::UNIMPLEMENTED, todo!()) */
  ZOS_ERR_CUDA = 3,           /* CUDA runtime / driver failure, or no device */
  ZOS_ERR_TYPE = 4,           /* operand descriptors do not match (CommandError::TYPE_ERR) */
  ZOS_ERR_STATE = 5,          /* call not valid in this state (StepError::ProgramEnd, StartError::MissingKey) */
  ZOS_ERR_OOM = 6
};

/* ---- texel vocabulary (image-canvas, re-exported by lib/zosimos/src/buffer.rs:2-6) ---- */
enum { /* Transfer; `as u32` discriminants, stage.frag:107-119 */
  ZOS_TRANSFER_BT709 = 0, ZOS_TRANSFER_BT470M = 1, ZOS_TRANSFER_BT601 = 2, ZOS_TRANSFER_SMPTE240 = 3,
  ZOS_TRANSFER_LINEAR = 4, ZOS_TRANSFER_SRGB = 5, ZOS_TRANSFER_BT2020_10BIT = 6, ZOS_TRANSFER_BT2020_12BIT = 7,
  ZOS_TRANSFER_SMPTE2084 = 8, ZOS_TRANSFER_BT2100PQ = 9, ZOS_TRANSFER_BT2100HLG = 10, ZOS_TRANSFER_LINEAR_SCENE = 11,
  ZOS_TRANSFER_LABLCH = 0x100 /* shaders/stage.rs:221-226 */
};
enum { /* SampleParts, shaders/stage.rs:74-96 */
  ZOS_PARTS_A = 0, ZOS_PARTS_R = 1, ZOS_PARTS_G = 2, ZOS_PARTS_B = 3, ZOS_PARTS_LUMA = 4, ZOS_PARTS_LUMAA = 5,
  ZOS_PARTS_RGB = 6, ZOS_PARTS_BGR = 7, ZOS_PARTS_RGBA = 8, ZOS_PARTS_RGBX = 9, ZOS_PARTS_BGRA = 10, ZOS_PARTS_BGRX = 11,
  ZOS_PARTS_ARGB = 12, ZOS_PARTS_XRGB = 13, ZOS_PARTS_ABGR = 14, ZOS_PARTS_XBGR = 15, ZOS_PARTS_YUV = 16,
  ZOS_PARTS_LAB = 17, ZOS_PARTS_LABA = 18, ZOS_PARTS_LCH = 19, ZOS_PARTS_LCHA = 20
};
enum { /* SampleBits, shaders/stage.rs:98-119 */
  ZOS_BITS_UINT8 = 0, ZOS_BITS_UINT332 = 1, ZOS_BITS_UINT233 = 2, ZOS_BITS_UINT16 = 3, ZOS_BITS_UINT4X4 = 4,
  ZOS_BITS_UINT_444 = 5, ZOS_BITS_UINT444_ = 6, ZOS_BITS_UINT565 = 7, ZOS_BITS_UINT8X2 = 8, ZOS_BITS_UINT8X3 = 9,
  ZOS_BITS_UINT8X4 = 10, ZOS_BITS_UINT16X2 = 11, ZOS_BITS_UINT16X3 = 12, ZOS_BITS_UINT16X4 = 13,
  ZOS_BITS_UINT2101010 = 14, ZOS_BITS_UINT1010102 = 15, /* 15 = the usual RGB10A2 (R in the low bits) */
  ZOS_BITS_UINT101010_ = 16, ZOS_BITS_UINT_101010 = 17, ZOS_BITS_FLOAT16X4 = 18, ZOS_BITS_FLOAT32X4 = 19
};
enum { ZOS_COLOR_RGB = 0, ZOS_COLOR_SCALARS = 1, ZOS_COLOR_OKLAB = 2, ZOS_COLOR_SRLAB2 = 3 }; /* image_canvas::Color */
enum { ZOS_PRIM_BT709 = 0, ZOS_PRIM_BT601_525 = 1, ZOS_PRIM_BT601_625 = 2, ZOS_PRIM_SMPTE240 = 3, ZOS_PRIM_BT2020 = 4, ZOS_PRIM_BT2100 = 5 };
enum { ZOS_WP_A = 0, ZOS_WP_B, ZOS_WP_C, ZOS_WP_D50, ZOS_WP_D55, ZOS_WP_D65, ZOS_WP_D75, ZOS_WP_E, ZOS_WP_F2, ZOS_WP_F7, ZOS_WP_F11 };
/* Block: the reference only lowers Block::Pixel (program.rs:794-938); the planar blocks are
 * additions of this backend */
enum { ZOS_BLOCK_PIXEL = 0, ZOS_BLOCK_YUV420_PLANAR = 1 /* I420: Y, U, V planes */, ZOS_BLOCK_YUV420_NV12 = 2 };
enum { ZOS_YUV_BT601 = 0, ZOS_YUV_BT709 = 1, ZOS_YUV_BT2020 = 2 }; /* luma coefficients Kr,Kb */

/* Replaces buffer::Descriptor {layout: ByteLayout, color: Color, texel: Texel}
 * (lib/zosimos/src/buffer.rs:14-30). */
typedef struct zos_desc {
  uint32_t width, height;
  uint64_t row_stride;   /* bytes between rows of plane 0 in the DEVICE buffer (256-aligned) */
  uint32_t texel_stride; /* bytes per texel (plane 0) */
  uint32_t block;        /* ZOS_BLOCK_* */
  uint32_t bits;         /* ZOS_BITS_* */
  uint32_t parts;        /* ZOS_PARTS_* */
  uint32_t color;        /* ZOS_COLOR_* */
  uint32_t transfer;     /* ZOS_TRANSFER_* (Color::Rgb / Color::Scalars) */
  uint32_t primaries;    /* ZOS_PRIM_* */
  uint32_t whitepoint;   /* ZOS_WP_* */
  uint32_t yuv_matrix;   /* ZOS_YUV_*            (planar blocks only) */
  uint32_t yuv_full_range;
  uint32_t chroma_filter; /* 0 nearest, 1 bilinear chroma upsampling */
  uint32_t reserved;
} zos_desc;

/* How a register's bytes turn into working values: the native-vs-staged decision of
 * ImageDescriptor::new (lib/zosimos/src/program.rs:781-946) plus this backend's float texels. */
enum { ZOS_STORAGE_STAGED = 0, ZOS_STORAGE_SRGB8 = 1, ZOS_STORAGE_UNORM8 = 2, ZOS_STORAGE_FLOAT = 3, ZOS_STORAGE_YUV420 = 4 };
typedef struct zos_texfmt { /* = the uvec4 of XyzParameter::serialize_std140, shaders/stage.rs:53-61 */
  uint32_t transfer, parts, bits, storage;
} zos_texfmt;

uint32_t zos_abi_version(void);
uint32_t zos_bits_bytes(uint32_t bits);                                   /* SampleBits::bytes() */
uint64_t zos_aligned_row_stride(uint32_t width, uint32_t texel_stride);   /* buffer.rs:121-134 */
zos_status zos_desc_texfmt(const zos_desc* desc, zos_texfmt* out);        /* program.rs:781-946 */
uint64_t zos_desc_device_bytes(const zos_desc* desc);                     /* all planes, buffer.rs:137-140 */

/* ---- context, device memory, host <-> device (replaces Pool's Gpu + ImageData::GpuBuffer,
 *      lib/zosimos/src/pool.rs:39-41,122-156, and Low::WriteImageToBuffer / Low::ReadBuffer,
 *      lib/zosimos/src/run.rs:1896-2041,2160-2276) ---- */
typedef struct zos_ctx zos_ctx;
typedef struct zos_buf zos_buf;

zos_status zos_ctx_create(int32_t device, zos_ctx** out);
void zos_ctx_destroy(zos_ctx* ctx);
const char* zos_last_error(const zos_ctx* ctx); /* ctx may be NULL: last error of a failed create */
int32_t zos_ctx_device(const zos_ctx* ctx);
void* zos_ctx_stream(const zos_ctx* ctx);       /* the cudaStream_t all launches of this ctx go to */
zos_status zos_sync(zos_ctx* ctx);             /* SyncPoint::block_on, run.rs:3019 */
/* SyncPoint::finish (run.rs:3041-3169, tests/async.rs): the non-blocking form.  *done = 1 when everything enqueued on the context has
 * finished (a kernel fault is reported like zos_sync does), 0 while work is still running; an asynchronous host polls this between
 * yields instead of blocking a thread in zos_sync. */
zos_status zos_poll(zos_ctx* ctx, int32_t* done);
/* Verification hook, host only (no context, no GPU): the rounding thresholds of the correctly rounded sRGB8
 * encoder (thr[k] = smallest f32 whose code is >= k; [0] = -inf, [256..259] = +inf) and the two bucket tables
 * the kernels derive from them (zosimos_b200/csrc/texel.cuh).  Arrays may be NULL; buckets holds up to 2048
 * entries.  The reference leaves this encode to the texture unit (program.rs:794-838). */
/* Host-only view of how k_affine_f16 sizes its staged source box (affine_f16.cu: the width, in texels, whose rows put the
 * taps of a 32-lane row on the fewest shared-memory bank conflicts for a source step of (step_x, step_y) texels per lane);
 * -1 for a width outside [2, 256].  For verification without a GPU. */
int32_t zos_affine_box_width(int32_t min_width, float step_x_per_lane, float step_y_per_lane);
zos_status zos_srgb_encoder_tables(float* thresholds260, uint32_t* buckets, uint32_t* n_buckets, uint32_t* buckets2, uint32_t* n_buckets2);
uint64_t zos_ctx_launch_count(const zos_ctx* ctx); /* kernels launched so far (bench: gpu_launches) */
/* debugging / parity switches; ZOS_CTX_NO_FAST_PATHS routes every launch through the generic kernels */
enum { ZOS_CTX_NO_FAST_PATHS = 1,
       ZOS_CTX_FRAME_FAST_ONLY = 2 /* frame pipeline: the one-role kernel (k_frame_fast) instead of the warp-specialised one: A/B and parity */ };
zos_status zos_ctx_set_flags(zos_ctx* ctx, uint32_t flags);

/* Device memory comes from a per-context arena of size-class free lists: zos_buf_free parks the block instead of
 * returning it to the driver and does NOT synchronise (work of a context is ordered on its one stream, so the next
 * owner's kernels run after the previous owner's); zos_buf_alloc takes a parked block of the class when there is one.
 * A relaunched program therefore allocates nothing after its first run -- the role of the pool cache behind
 * Environment::recover_buffers / Retire::retire_buffers (lib/zosimos/src/run.rs:1312-1347, 2876-2942; pool.rs:93-99).
 * A buffer that another context (stream) still uses must not be freed before that context was synchronised. */
zos_status zos_buf_alloc(zos_ctx* ctx, uint64_t bytes, zos_buf** out);
void zos_buf_free(zos_ctx* ctx, zos_buf* buf);
typedef struct zos_arena_stats {
  uint64_t device_allocs;   /* cudaMalloc calls made for buffers so far */
  uint64_t reuses;          /* allocations served from a parked block */
  uint64_t bytes_reserved;  /* device memory held by the arena (in use + parked) */
  uint64_t bytes_in_use;
  uint64_t bytes_parked;
} zos_arena_stats;
zos_status zos_ctx_arena_stats(const zos_ctx* ctx, zos_arena_stats* out);
/* Pool::clear_cache (pool.rs:450-455): hand every parked block back to the driver (synchronises the context) */
zos_status zos_ctx_arena_trim(zos_ctx* ctx);
void* zos_buf_ptr(const zos_buf* buf);
uint64_t zos_buf_size(const zos_buf* buf);
/* pinned host staging memory (the map_write / map_read buffers of encoder.rs:574-616) */
zos_status zos_host_alloc(zos_ctx* ctx, uint64_t bytes, void** out);
void zos_host_free(zos_ctx* ctx, void* ptr);
/* rows x row_bytes from tight/pitched host rows into the pitched device buffer and back
 * (copy_host_to_buffer, run.rs:3282-3309; the ReadBuffer row loop, run.rs:2265-2270).
 * Asynchronous on the ctx stream when `host` is pinned. */
zos_status zos_buf_upload(zos_ctx* ctx, zos_buf* dst, uint64_t dst_offset, uint64_t dst_pitch, const void* host,
                          uint64_t host_pitch, uint64_t row_bytes, uint64_t rows);
zos_status zos_buf_download(zos_ctx* ctx, const zos_buf* src, uint64_t src_offset, uint64_t src_pitch, void* host,
                            uint64_t host_pitch, uint64_t row_bytes, uint64_t rows);
zos_status zos_buf_copy(zos_ctx* ctx, zos_buf* dst, uint64_t dst_offset, const zos_buf* src, uint64_t src_offset,
                        uint64_t bytes); /* High::Copy (transmute / from_buffer), program.rs:1534-1576 */
zos_status zos_buf_fill(zos_ctx* ctx, zos_buf* dst, uint64_t offset, uint64_t bytes, uint8_t value);

/* ---- images: a descriptor + where its planes live ---- */
typedef struct zos_image {
  zos_desc desc;
  void* data;            /* device pointer, plane 0 (pixels, or Y) */
  void* plane1;          /* U (I420) or interleaved UV (NV12); NULL for ZOS_BLOCK_PIXEL */
  void* plane2;          /* V (I420) */
  uint64_t chroma_stride; /* bytes between chroma rows */
  uint64_t batch_stride;  /* bytes between consecutive frames of plane 0 (0 if batch == 1) */
  uint64_t chroma_batch_stride;
} zos_image;

/* whole-image transfers between TIGHT host rows and a (pitched) device image, frame `frame` of a
 * batch.  Host layout: plane 0 rows of width*texel_stride bytes; planar 4:2:0 then U and V rows
 * (I420) or the interleaved UV rows (NV12).  Asynchronous on the ctx stream when `host` is pinned. */
zos_status zos_image_upload(zos_ctx* ctx, const zos_image* dst, uint32_t frame, const void* host);
zos_status zos_image_download(zos_ctx* ctx, const zos_image* src, uint32_t frame, void* host);

/* ---- per-pixel steps fused into a kernel between unpack and pack.  Parameter layouts are the
 *      reference's uniform blocks with matrices row-major. ---- */
enum {
  ZOS_STEP_MATRIX = 1,     /* linear.frag:12-17            m = 3x3                       */
  ZOS_STEP_OKLAB_ENC = 2,  /* oklab.frag:34-47             m = rgb -> xyz                */
  ZOS_STEP_OKLAB_DEC = 3,  /* oklab.frag:50-64             m = xyz -> rgb                */
  ZOS_STEP_SRLAB2_ENC = 4, /* srlab2.frag:36-57            m = rgb -> xyz                */
  ZOS_STEP_SRLAB2_DEC = 5, /* srlab2.frag:60-85            m = xyz -> rgb, v = whitepoint */
  ZOS_STEP_REQUANT = 6,    /* store to + reload from a declared register: fmt            */
  ZOS_STEP_INJECT = 7,     /* inject.frag:20-25 (second operand = `aux` image)  v=mix,m[0..3]=color */
  ZOS_STEP_F16 = 8         /* round through an Rgba16Float texture                        */
};
typedef struct zos_step {
  uint32_t kind;
  zos_texfmt fmt; /* ZOS_STEP_REQUANT */
  float m[9];
  float v[4];
} zos_step;
#define ZOS_MAX_STEPS 8

/* unpack(src) -> steps -> pack(dst), one pass, `batch` frames.  Replaces the decode pass, the
 * PaintFullScreen draw(s) and the encode pass of color_convert / chromatic_adaptation / extract
 * (command.rs:2527-2597; program.rs:1475-1533). */
zos_status zos_pixel_chain(zos_ctx* ctx, const zos_image* src, const zos_image* dst, const zos_step* steps,
                           uint32_t nsteps, uint32_t batch);

/* ---- two-layer composition and resampling ---- */
enum { ZOS_SAMPLE_NEAREST = 0, ZOS_SAMPLE_BILINEAR = 1 }; /* AffineSample, command.rs:339-352 */
enum { /* how `above` texels land on the destination */
  ZOS_MAP_RECT = 0,   /* QuadTarget::Rect: selection sel[] stretched over target tgt[] (inscribe, crop,
                         blend; program.rs:1908-1920) -- exact rational nearest indexing */
  ZOS_MAP_AFFINE = 1, /* QuadTarget::Absolute: inv[] = inverse of Affine.transformation (program.rs:1898-1906) */
  ZOS_MAP_GRID8 = 2,  /* CommandBuffer::resize as the reference does it: an RGBA8 coordinate grid and a
                         palette lookup, coordinates truncated to 8 bits (command.rs:1675-1702) */
  ZOS_MAP_SCALE = 3   /* exact resize: tgt[] = full destination, half-pixel centres */
};
enum { /* what happens where `above` lands */
  ZOS_BLEND_OVERWRITE = -1, /* blend: None (encoder.rs:1493): RGBA replaced -- inscribe / affine */
  ZOS_BLEND_CLEAR = 0, ZOS_BLEND_SRC = 1, ZOS_BLEND_DST = 2, ZOS_BLEND_SRC_OVER = 3, ZOS_BLEND_DST_OVER = 4,
  ZOS_BLEND_SRC_IN = 5, ZOS_BLEND_DST_IN = 6, ZOS_BLEND_SRC_OUT = 7, ZOS_BLEND_DST_OUT = 8, ZOS_BLEND_SRC_ATOP = 9,
  ZOS_BLEND_DST_ATOP = 10, ZOS_BLEND_XOR = 11, /* Porter-Duff, linear light (Blend::Alpha = SRC_OVER) */
  ZOS_BLEND_INJECT = 12 /* inject.frag:20-25: mix(below, vec4(dot(above, inject_color)), inject_mix) */
};
typedef struct zos_compose_params {
  int32_t map;       /* ZOS_MAP_* */
  int32_t sampling;  /* ZOS_SAMPLE_* */
  int32_t blend;     /* ZOS_BLEND_* */
  int32_t use_tma;   /* 0 = direct loads, 1 = stage source tiles through TMA into shared memory when possible */
  int32_t sel[4];    /* x, y, w, h in `above` texels (ZOS_MAP_RECT) */
  int32_t tgt[4];    /* x, y, w, h in destination pixels (ZOS_MAP_RECT / ZOS_MAP_SCALE) */
  float inv[9];      /* row-major inverse affine (ZOS_MAP_AFFINE) */
  /* sharding one large image over GPUs by row bands / tiles (all zero = whole images): dst is the window of
   * the full destination starting at dst_origin, `above` the window of the full src_full[0] x src_full[1]
   * source starting at src_origin.  sel / tgt / inv stay in FULL-image coordinates, so a windowed launch
   * writes exactly the bytes of the whole-image launch. */
  int32_t dst_origin[2], src_origin[2], src_full[2];
  float inject_mix[4], inject_color[4]; /* ZOS_BLEND_INJECT (shaders/inject.rs:28-31) */
  uint32_t n_src_steps, n_dst_steps;
  zos_step src_steps[ZOS_MAX_STEPS]; /* applied to every `above` tap after unpack */
  zos_step dst_steps[ZOS_MAX_STEPS]; /* applied to the composed value before pack */
} zos_compose_params;

/* dst = pack(dst_steps(compose(unpack(below), sample(src_steps(unpack(above)))))).  `below` may be
 * NULL (pure resample; uncovered pixels get the Target::Discard colour (0,0,1,1), program.rs:1494-1506).
 * Replaces the two PaintToSelection passes of inscribe / affine (command.rs:2642-2740), resize's
 * bilinear + palette passes (command.rs:1675-1702) and implements blend (command.rs:1510-1519). */
zos_status zos_compose(zos_ctx* ctx, const zos_image* below, const zos_image* above, const zos_image* dst,
                       const zos_compose_params* params, uint32_t batch);

/* Generators (ConstructOp::{Solid,Bilinear,...}, command.rs:1524-1633; DrawInto without operands): kind + 24 floats
 * (unused ones ignored).
 *   BILINEAR      u_min,u_max,v_min,v_max,uv_min,uv_max (6 x vec4; bilinear.frag:14-20, shaders/bilinear.rs:34-45)
 *   SOLID         colour (solid_rgb.frag)
 *   NORMAL2D      expectation[2], covariance_inverse[4] row major, pseudo_determinant (distribution_normal2d.frag:43-55)
 *   FRACTAL_NOISE initial_scale[2], amplitude, damping, num_octaves (fractal_noise.frag; shaders/fractal_noise.rs:21-49)
 * In a zos_op of kind ZOS_OP_GENERATE the generator kind travels in compose.map. */
enum { ZOS_GEN_BILINEAR = 0, ZOS_GEN_SOLID = 1, ZOS_GEN_NORMAL2D = 2, ZOS_GEN_FRACTAL_NOISE = 3 };
zos_status zos_generate(zos_ctx* ctx, const zos_image* dst, uint32_t kind, const float* p, uint32_t batch);
/* WithBuffer (command.rs:1963-2060, tests/buffer.rs:121-149): the 24-float parameter block is read from device memory */
zos_status zos_generate_from_buffer(zos_ctx* ctx, const zos_image* dst, uint32_t kind, const zos_buf* params, uint64_t offset, uint32_t batch);
zos_status zos_generate_bilinear(zos_ctx* ctx, const zos_image* dst, const float* p, uint32_t batch);
zos_status zos_generate_solid(zos_ctx* ctx, const zos_image* dst, const float* color /* 4 floats */, uint32_t batch);
/* box3.frag:16-52 (derivative, command.rs:1493-1508): m = 3x3 weights, row-major [dy+1][dx+1] */
zos_status zos_box3(zos_ctx* ctx, const zos_image* src, const zos_image* dst, const float* m, uint32_t batch);
/* palette.frag:21-32 (command.rs:1442-1485): dst = pal @ (xc . idx, yc . idx) */
zos_status zos_palette(zos_ctx* ctx, const zos_image* pal, const zos_image* idx, const zos_image* dst, const float* xc,
                       const float* yc, uint32_t batch);

/* ---- programs: the High instruction stream (lib/zosimos/src/program.rs:89-139) ---- */
typedef struct zos_program zos_program;
enum {
  ZOS_OP_INPUT = 1,        /* High::Input(reg)                       dst                           */
  ZOS_OP_OUTPUT = 2,       /* High::Output{src,dst}                  src[0]                        */
  ZOS_OP_PIXEL = 3,        /* PushOperand + DrawInto{PaintFullScreen} src[0] -> dst, steps          */
  ZOS_OP_COMPOSE = 4,      /* the 2x PaintToSelection pattern        src[0]=below src[1]=above      */
  ZOS_OP_COPY = 5,         /* High::Copy (transmute)                 src[0] -> dst                  */
  ZOS_OP_GENERATE = 6,     /* DrawInto without operands (bilinear / solid)  gen[24]                 */
  ZOS_OP_BOX3 = 7,
  ZOS_OP_PALETTE = 8,      /* src[0]=palette src[1]=indices, compose.inv[0..7] = xc,yc              */
  /* byte buffers in device memory (command.rs:1777-1803 buffer_init / buffer_zero, :937-968 from_buffer,
   * WithBuffer :1963-2060; tests/buffer.rs).  A buffer register has no texel: desc is ignored. */
  ZOS_OP_BUFFER_INIT = 9,  /* dst = buffer register of data_len bytes, filled from `data` (NULL = zeroed)   */
  ZOS_OP_FROM_BUFFER = 10, /* src[0] = buffer register holding the image in the ALIGNED device layout
                              (row_stride = zos_aligned_row_stride) -> dst image register (High::Copy)       */
  ZOS_OP_DYNAMIC = 11      /* user operator (command.rs:2933-3060 {construct,unary,binary}_dynamic): `source` = CUDA C
                              defining zos_shade (see zos_dynamic_create), data / data_len = its parameter bytes,
                              src[0], src[1] = operands or -1, desc = the image it produces                   */
  /* ZOS_OP_GENERATE with src[0] = a buffer register: the parameter block is read from that buffer at run time */
};
typedef struct zos_op {
  uint32_t kind;
  int32_t src[2]; /* register indices, -1 = unused */
  int32_t dst;
  zos_desc desc;  /* descriptor of dst (Input: of the bound image) */
  uint32_t nsteps;
  zos_step steps[ZOS_MAX_STEPS];
  zos_compose_params compose;
  float gen[24];
  uint32_t knob; /* 0 = none, else 1-based knob id whose bytes overwrite this op's parameter block */
  int32_t reg;   /* this op's own register number (Register(idx), command.rs:31); equals dst except for Output ops */
  const void* data;  /* ZOS_OP_BUFFER_INIT: initial bytes (copied by zos_program_create), or NULL; ZOS_OP_DYNAMIC: parameter bytes */
  uint64_t data_len; /* ZOS_OP_BUFFER_INIT: size of the buffer in bytes; ZOS_OP_DYNAMIC: size of the parameter block */
  const char* source; /* ZOS_OP_DYNAMIC: NUL-terminated CUDA C source (copied by zos_program_create) */
} zos_op;
enum { ZOS_FUSE_EXACT = 0 /* every declared register is quantised like the reference, in registers */,
       ZOS_FUSE_WIDE = 1 /* fused intermediates stay f32 */,
       ZOS_FUSE_NONE = 2 /* one kernel per op, intermediates materialised (debug / parity) */ };

/* Program::lower_to (program.rs:1304-1423): plan buffers, fuse, build the kernel schedule */
zos_status zos_program_create(zos_ctx* ctx, const zos_op* ops, uint32_t nops, uint32_t fuse_mode, uint32_t batch,
                              zos_program** out);
void zos_program_destroy(zos_program* prog);
/* Environment::bind / bind_output (run.rs:1171-1244): registers of Input ops and the src of Output ops */
zos_status zos_program_bind(zos_program* prog, int32_t reg, const zos_image* image);
/* undo a binding before the program is launched again with other images: an input is unbound again, an output register
 * gets its storage from the program again */
zos_status zos_program_unbind(zos_program* prog, int32_t reg);
/* Environment::knob (run.rs:1292-1306) */
zos_status zos_program_set_knob(zos_program* prog, uint32_t knob, const void* data, uint64_t len);
/* every knob back to the parameter block the program was planned with: knobs belong to one Environment (run.rs:1292-1306),
 * so a cached program is reset before the next environment's knobs are applied */
zos_status zos_program_reset_knobs(zos_program* prog);
/* Executable::launch + Execution::step (run.rs:1016,1389): step launches up to max_kernels kernels */
zos_status zos_program_launch(zos_program* prog);
zos_status zos_program_step(zos_program* prog, uint32_t max_kernels, int32_t* still_running);
/* User operators: the CUDA analogue of the reference's plugin interface `trait ShaderCommand` (command/dynamic.rs:7-60,
 * SPIR-V fragment shaders there).  `cuda_source` defines
 *     __device__ float4 zos_shade(float2 uv, const unsigned char* params, zos_tex in0, zos_tex in1);
 * evaluated per destination pixel at uv = pixel centre / size; in0.fetch(uv) / in0.at(x, y) read the operands' working
 * values (nearest texel).  Compiled with NVRTC for sm_100a on first use, cached per context by source text; the
 * compiler log is in zos_last_error on failure.  A zos_dynamic belongs to its context (freed by zos_ctx_destroy). */
typedef struct zos_dynamic zos_dynamic;
zos_status zos_dynamic_create(zos_ctx* ctx, const char* cuda_source, zos_dynamic** out);
zos_status zos_dynamic_launch(zos_ctx* ctx, zos_dynamic* dyn, const zos_image* dst, const zos_image* in0, const zos_image* in1,
                              const void* params, uint64_t params_len);

/* Executable reuse (run.rs:1016,1283-1347; Readme.md "re-use of the pipeline"; tests/loop.rs, tests/knobs.rs):
 * run the whole schedule again, same plan, possibly other bindings / knob values.  ZOS_RUN_GRAPH relaunches
 * the schedule as one CUDA graph (captured on the second run, re-captured after zos_program_bind /
 * zos_program_set_knob).  zos_program_graph_launches counts the runs that went through the graph. */
enum { ZOS_RUN_EAGER = 0, ZOS_RUN_GRAPH = 1 };
zos_status zos_program_run(zos_program* prog, uint32_t flags);
uint64_t zos_program_graph_launches(const zos_program* prog);
uint32_t zos_program_kernel_count(const zos_program* prog);
/* Resource recovery between launches (Environment::recover_buffers, run.rs:1312-1347; Retire::retire_buffers, run.rs:2876-2942;
 * tests/util.rs:95-118).  A program provides storage for every register that is neither a bound input nor a bound output.
 * zos_program_release_buffers parks that storage in the context's arena (any program may take it from there);
 * zos_program_recover_buffers takes it back -- bytes_reused came from parked blocks, bytes_allocated needed cudaMalloc --
 * and zos_program_launch / zos_program_run do so implicitly.  Either pointer may be NULL. */
typedef struct zos_program_stats {
  uint32_t kernels;       /* launches per run of the fused schedule */
  uint32_t temp_buffers;  /* registers the program provides storage for */
  uint64_t temp_bytes;
  uint32_t released;      /* 1 while that storage is parked in the arena */
  uint32_t reserved;
  uint64_t runs, graph_launches;
} zos_program_stats;
zos_status zos_program_release_buffers(zos_program* prog, uint64_t* bytes, uint32_t* count);
zos_status zos_program_recover_buffers(zos_program* prog, uint64_t* bytes_reused, uint64_t* bytes_allocated);
zos_status zos_program_resources(const zos_program* prog, zos_program_stats* out);
/* fills descriptor + device location of a register the program allocated itself (outputs not bound) */
zos_status zos_program_register_image(const zos_program* prog, int32_t reg, zos_image* out);

/* ---- several GPUs.  The path shards into independent units -- frames of a batch, row bands of one large
 * image -- with no exchange step; the reference always uses the first device of its pool (pool.rs:227-230; "If multi-device
 * then this should become a set", run.rs:416-420).  What a host needs across devices:
 *   zos_multi_launch   one host thread starts the programs of several contexts (zos_program_run each, ZOS_RUN_* flags)
 *                      without waiting; zos_multi_sync waits for all of them and reports the first error
 *   zos_gather_peer    contexts of ONE process: shard i (srcs[i] + src_offsets[i], bytes[i]) is copied over NVLink to
 *                      dst + dst_offsets[i]; each copy is ordered after the work already enqueued on its source context,
 *                      and the destination context's stream waits for all of them (no host synchronisation)
 *   zos_comm_* / zos_gather_nccl   one process per GPU: an NCCL communicator bound to the context (NCCL is loaded at run
 *                      time, libnccl.so.2 or $ZOS_NCCL_LIBRARY; ZOS_ERR_UNSUPPORTED if absent).  Rank 0 makes the 128-byte
 *                      id with zos_comm_unique_id and hands it to the other ranks by any host channel.  zos_gather_nccl:
 *                      every rank contributes shard_bytes[rank] bytes from send + send_off; rank `root` (or every rank when
 *                      root < 0) receives shard r at recv + recv_offsets[r].  Enqueued on the context's stream. */
typedef struct zos_comm zos_comm;
zos_status zos_multi_launch(zos_program* const* progs, uint32_t n, uint32_t flags);
zos_status zos_multi_sync(zos_ctx* const* ctxs, uint32_t n);
zos_status zos_gather_peer(zos_ctx* dst_ctx, zos_buf* dst, const uint64_t* dst_offsets, zos_ctx* const* src_ctxs,
                           const zos_buf* const* srcs, const uint64_t* src_offsets, const uint64_t* bytes, uint32_t n);
zos_status zos_comm_unique_id(uint8_t* id128);
zos_status zos_comm_create(zos_ctx* ctx, const uint8_t* id128, uint32_t rank, uint32_t world, zos_comm** out);
void zos_comm_destroy(zos_comm* comm);
int32_t zos_comm_nccl_version(void); /* e.g. 22809; 0 when NCCL cannot be loaded */
zos_status zos_gather_nccl(zos_comm* comm, const zos_buf* send, uint64_t send_off, zos_buf* recv, const uint64_t* recv_offsets,
                           const uint64_t* shard_bytes, int32_t root);

#ifdef __cplusplus
}
#endif
#endif /* ZOSIMOS_CUDA_H */
