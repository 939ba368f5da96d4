/* zosimos_host.h -- C view of the C++ host layer that mirrors the reference's operator API
 * (`CommandBuffer` and `Linker`, lib/zosimos/src/command.rs) above the device C-ABI of
 * zosimos_cuda.h.  A Rust maintainer keeps the reference's own CommandBuffer/Linker and binds at
 * zos_program_create; this layer exists for C/C++/Python callers (the reference's toolchain is not
 * available in the build image, so the parity tests drive this mirror).  Same names, argument
 * meaning and error behaviour as the reference; every function cites the method it mirrors.
 */
#ifndef ZOSIMOS_HOST_H
#define ZOSIMOS_HOST_H

#include "zosimos_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* CommandErrorKind, lib/zosimos/src/command.rs:3612-3647 */
enum {
  ZOSH_OK = 0,
  ZOSH_ERR_BAD_DESCRIPTOR = 1,    /* CommandErrorKind::BadDescriptor */
  ZOSH_ERR_CONFLICTING_TYPES = 2, /* CommandErrorKind::ConflictingTypes */
  ZOSH_ERR_TYPE = 3,              /* CommandError::TYPE_ERR (GenericTypeError) */
  ZOSH_ERR_OTHER = 4,             /* CommandError::OTHER (also BAD_REGISTER / INVALID_CALL) */
  ZOSH_ERR_UNIMPLEMENTED = 5,     /* CommandError::UNIMPLEMENTED / CompileError::NotYetImplemented */
  ZOSH_ERR_CONCRETE_REQUIRED = 6  /* CommandErrorKind::ConcreteDescriptorRequired */
};
enum { ZOSH_ADAPT_BRADFORD_VONKRIES = 0, ZOSH_ADAPT_VONKRIES = 1, ZOSH_ADAPT_XYZ = 2, ZOSH_ADAPT_BRADFORD_NONLINEAR = 3 };
enum { ZOSH_RESIZE_REFERENCE = 0 /* 8-bit grid + palette, command.rs:1675-1702 */, ZOSH_RESIZE_NEAREST = 1, ZOSH_RESIZE_BILINEAR = 2 };
enum { ZOSH_DERIV_PREWITT = 0, ZOSH_DERIV_SOBEL = 1, ZOSH_DERIV_SCHARR3 = 2, ZOSH_DERIV_SCHARR3_TO_4BIT = 3, ZOSH_DERIV_SCHARR3_TO_8BIT = 4,
       ZOSH_DERIV_ROBERTS = 5 /* NotYetImplemented, command.rs:3402-3408 */ };

typedef struct zosh_cb zosh_cb;           /* command::CommandBuffer */
typedef struct zosh_program zosh_program; /* program::Program (the linked High stream) */
typedef struct zosh_signature zosh_signature; /* command::CommandSignature (the callee travels inside it) */
typedef struct zosh_rect { uint32_t x, y, max_x, max_y; } zosh_rect; /* command::Rectangle, command.rs:311-317 */

const char* zosh_last_error(void); /* message of the last failure on this thread */

/* colour science used by the op builder (image-canvas Primaries::to_xyz_row_matrix, palette
 * TransformMatrix; call sites command.rs:1023-1073, 3281-3338).  Row-major 3x3, f32. */
int32_t zosh_to_xyz_matrix(uint32_t primaries, uint32_t whitepoint, float out[9]);
int32_t zosh_adaptation_matrix(uint32_t method, uint32_t src_wp, uint32_t dst_wp, float out[9]);
int32_t zosh_whitepoint_xyz(uint32_t whitepoint, float out[3]);
/* Affine::{new,scale,rotate,shift}: left multiplication in f32, command.rs:3421-3485 */
void zosh_affine_identity(float m[9]);
void zosh_affine_scale(float m[9], float x, float y);
void zosh_affine_rotate(float m[9], float rad);
void zosh_affine_shift(float m[9], float x, float y);
/* Rectangle::normalize with the reference's max_y = y + width (command.rs:3536-3543) */
zosh_rect zosh_rect_normalize(zosh_rect r);

zosh_cb* zosh_cb_new(void);
void zosh_cb_free(zosh_cb* cb);
/* every builder returns ZOSH_OK and the new register in *reg, or an error kind */
int32_t zosh_cb_input(zosh_cb* cb, const zos_desc* desc, int32_t* reg);                               /* command.rs:743 */
int32_t zosh_cb_output(zosh_cb* cb, int32_t src, int32_t* reg);                                      /* command.rs:1707 */
uint32_t zosh_cb_num_ops(const zosh_cb* cb);                                                         /* number of operations pushed so far */
int32_t zosh_cb_describe(const zosh_cb* cb, int32_t reg, zos_desc* out);                             /* describe_reg */
int32_t zosh_cb_color_convert(zosh_cb* cb, int32_t src, const zos_desc* color_and_texel, int32_t* reg); /* command.rs:986 */
int32_t zosh_cb_chromatic_adaptation(zosh_cb* cb, int32_t src, uint32_t method, uint32_t target_wp, int32_t* reg); /* :1112 */
int32_t zosh_cb_inscribe(zosh_cb* cb, int32_t below, zosh_rect rect, int32_t above, int32_t* reg);   /* command.rs:1177 */
int32_t zosh_cb_crop(zosh_cb* cb, int32_t src, zosh_rect rect, int32_t* reg);                        /* command.rs:971 */
int32_t zosh_cb_affine(zosh_cb* cb, int32_t below, const float m[9], uint32_t sampling, int32_t above, int32_t* reg); /* :1636 */
int32_t zosh_cb_resize(zosh_cb* cb, int32_t below, uint32_t w, uint32_t h, uint32_t mode, int32_t* reg); /* command.rs:1675 */
int32_t zosh_cb_blend(zosh_cb* cb, int32_t below, zosh_rect rect, int32_t above, int32_t mode, int32_t* reg); /* command.rs:1510 */
int32_t zosh_cb_transmute(zosh_cb* cb, int32_t src, const zos_desc* target, int32_t* reg);            /* command.rs:1276 */
int32_t zosh_cb_bilinear(zosh_cb* cb, const zos_desc* desc, const float p[24], int32_t* reg);         /* command.rs:1615 */
int32_t zosh_cb_solid_rgba(zosh_cb* cb, const zos_desc* desc, const float color[4], int32_t* reg);    /* command.rs:1524 */
/* generators of the std library next to `bilinear` (command.rs distribution_normal2d / distribution_fractal_noise;
 * shaders/distribution_normal2d.rs:25-100, shaders/fractal_noise.rs:21-49).  params as in zos_generate. */
void zosh_normal2d_with_diagonal(float var0, float var1, float out[7]);
void zosh_normal2d_with_direction(float x, float y, float out[7]);
void zosh_fractal_noise_with_octaves(uint32_t octaves, float out[5]);
void zosh_fractal_noise_set_damping(float params[5], float damping);
int32_t zosh_cb_distribution_normal2d(zosh_cb* cb, const zos_desc* desc, const float params[7], int32_t* reg);
int32_t zosh_cb_distribution_fractal_noise(zosh_cb* cb, const zos_desc* desc, const float params[5], int32_t* reg);
int32_t zosh_cb_derivative(zosh_cb* cb, int32_t src, uint32_t method, uint32_t height_direction, int32_t* reg); /* :1493 */
int32_t zosh_cb_palette(zosh_cb* cb, int32_t palette, int32_t indices, const float xc[4], const float yc[4], int32_t* reg); /* :1442 */
/* channel: 0 R, 1 G, 2 B, 3 Alpha (ColorChannel); extract = full copy whose destination texel keeps one channel */
int32_t zosh_cb_extract(zosh_cb* cb, int32_t src, uint32_t channel, int32_t* reg);                       /* command.rs:1221 */
int32_t zosh_cb_inject(zosh_cb* cb, int32_t below, uint32_t channel, int32_t above, int32_t* reg);      /* command.rs:1360 */
/* byte buffers in device memory and what consumes them (tests/buffer.rs) */
int32_t zosh_cb_buffer_init(zosh_cb* cb, const void* data, uint64_t len, int32_t* reg);                 /* command.rs:1777 (knob: :1938) */
int32_t zosh_cb_buffer_zero(zosh_cb* cb, uint64_t len, int32_t* reg);                                   /* command.rs:1793 */
int32_t zosh_cb_buffer_size(const zosh_cb* cb, int32_t reg, uint64_t* out);                             /* RegisterDescription::Buffer */
int32_t zosh_cb_from_buffer(zosh_cb* cb, int32_t buffer, const zos_desc* desc, int32_t* reg);           /* command.rs:937-968 */
int32_t zosh_cb_with_buffer_bilinear(zosh_cb* cb, int32_t buffer, const zos_desc* desc, int32_t* reg);  /* command.rs:1963-2060 + bilinear */
/* user operators: construct_dynamic (no operand), unary_dynamic (src0), binary_dynamic (src0, src1), command.rs:2933-3060.
 * `cuda_source` is the plugin (see zos_dynamic_create), `desc` the descriptor ShaderCommand::data returns, params its data. */
int32_t zosh_cb_dynamic(zosh_cb* cb, int32_t src0, int32_t src1, const char* cuda_source, const zos_desc* desc, const void* params,
                        uint64_t params_len, int32_t* reg);
int32_t zosh_cb_with_knob(zosh_cb* cb);  /* the NEXT operation gets a knob; returns its 1-based id (command.rs:1865-1874) */

/* Functions and generics (command.rs:856-922, 2821-2869; tests/generic.rs).  A command buffer that declares a generic is a
 * TEMPLATE: its builders record their calls (registers = positions in the record; zosh_cb_describe has nothing to answer
 * until types are bound) and zosh_cb_invoke replays the record into the caller with the generic inputs replaced by the
 * argument registers -- monomorphisation by inlining, which the reference does at link time (command.rs:2083-2185).
 * Type errors of the callee under the bound types therefore surface at invoke, and leave the caller unchanged. */
int32_t zosh_cb_generic(zosh_cb* cb, int32_t* var);                                                      /* command.rs:856-870 */
int32_t zosh_cb_input_generic(zosh_cb* cb, int32_t var, int32_t* reg);                                   /* command.rs:872-884 */
int32_t zosh_cb_computed_signature(const zosh_cb* cb, zosh_signature** out);                             /* command.rs:886-905 */
void zosh_signature_free(zosh_signature* sig);
uint32_t zosh_signature_num_generics(const zosh_signature* sig);
uint32_t zosh_signature_num_inputs(const zosh_signature* sig);
uint32_t zosh_signature_num_outputs(const zosh_signature* sig);
int32_t zosh_cb_function(zosh_cb* cb, const zosh_signature* sig, int32_t* function);                     /* command.rs:907-922 */
uint32_t zosh_cb_num_functions(const zosh_cb* cb);
/* Inside a template, a generic of the callee can be bound to one of the template's own generics: pass a zos_desc whose
 * `reserved` field is ZOSH_GENERIC_VAR | var (everything else ignored).  Such calls are recorded and inlined, with their
 * types checked, when the enclosing template is itself invoked. */
#define ZOSH_GENERIC_VAR 0x80000000u
/* InvocationArguments{generics, arguments}; results[] receives the registers of the callee's outputs, in order.
 * ZOSH_ERR_TYPE = CommandError::INVALID_CALL (count or type mismatch), ZOSH_ERR_OTHER = BAD_REGISTER. */
int32_t zosh_cb_invoke(zosh_cb* cb, int32_t function, const zos_desc* generics, uint32_t num_generics, const int32_t* arguments,
                       uint32_t num_arguments, int32_t* results, uint32_t results_cap, uint32_t* num_results); /* command.rs:2821-2869 */
/* Linker::link (command.rs:2083-2185): program 0 is `main`, program k >= 1 is functions[k - 1]; the link tables of all programs
 * are concatenated in `links` (links_per_program[p] entries for program p), entry f of a table names the program that function
 * variable f calls.  Calls were inlined by zosh_cb_invoke, so linking verifies the wiring and compiles `main`.
 * `tys` binds the generics of a generic entry point (main itself a template; num_tys = its number of generics, else 0): the
 * compiled stream is main's monomorphic copy, and zosh_program_register translates main's registers (inputs to bind, outputs
 * to retire, knobs keep their ids) into the program's; for a non-generic main it is the identity.  -1 = no such register. */
int32_t zosh_link(const zosh_cb* main_cb, const zos_desc* tys, uint32_t num_tys, const zosh_cb* const* functions, uint32_t num_functions,
                  const uint32_t* links, const uint32_t* links_per_program, zosh_program** out);
int32_t zosh_program_register(const zosh_program* p, int32_t reg);
/* Executable::query_knob for RegisterKnob{link_idx, register} (command.rs:701-705, 2134-2145): the Knob id the linker
 * assigned (ids count in emission order, knobs inside invoked functions included), 0 = none.  link_idx 0 = main's own
 * registers (the template's registers for a generic entry point), k >= 1 = registers of functions[k - 1]; where a
 * function was instantiated several times the last instantiation answers, like the reference's map insert. */
uint32_t zosh_program_knob(const zosh_program* p, uint32_t link_idx, int32_t reg);

/* Linker::compile (command.rs:2069): liveness + emission of the High-like op list */
int32_t zosh_compile(const zosh_cb* cb, zosh_program** out);
void zosh_program_free(zosh_program* p);
uint32_t zosh_program_num_ops(const zosh_program* p);
const zos_op* zosh_program_ops(const zosh_program* p);
/* Program::lower_to (program.rs:1304): hand the stream to the device planner */
zos_status zosh_program_lower(const zosh_program* p, zos_ctx* ctx, uint32_t fuse_mode, uint32_t batch, zos_program** out);

#ifdef __cplusplus
}
#endif
#endif
