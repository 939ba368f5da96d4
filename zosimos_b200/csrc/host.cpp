// host.cpp -- C++ mirror of the reference's operator API above the device C-ABI:
// command::CommandBuffer (typed SSA op builder with the reference's checks and error kinds,
// lib/zosimos/src/command.rs:743-1740), Linker::compile (liveness + High emission,
// command.rs:2069-2893) and the colour science its builders call (image-canvas / palette; see
// zosimos_host.h).  Everything here is host arithmetic; the output is a zos_op[] stream for
// zos_program_create.
#include <math.h>
#include <string.h>

#include <memory>
#include <string>
#include <vector>

#include "../../include/zosimos_host.h"

namespace {

thread_local std::string g_err;
int32_t err(int32_t kind, const char* msg) { g_err = msg; return kind; }

// ---- 3x3 algebra in double, fixed evaluation order (the oracle mirrors these formulas) ----
void inv3(const double* m, double* o) {
  double a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
  double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  double det = a * A + b * B + c * C;
  o[0] = A / det; o[1] = -(b * i - c * h) / det; o[2] = (b * f - c * e) / det;
  o[3] = B / det; o[4] = (a * i - c * g) / det; o[5] = -(a * f - c * d) / det;
  o[6] = C / det; o[7] = -(a * h - b * g) / det; o[8] = (a * e - b * d) / det;
}
void mul3(const double* a, const double* b, double* o) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c] + a[3 * r + 2] * b[6 + c];
}
void to_f32(const double* m, float* o, int n = 9) { for (int i = 0; i < n; i++) o[i] = (float)m[i]; }

// CIE xy chromaticities of the primaries (ITU-R BT.601/709/2020, SMPTE 240M), image-canvas `Primaries`
const double PRIM[6][6] = {
    {0.64, 0.33, 0.30, 0.60, 0.15, 0.06},      // Bt709
    {0.630, 0.340, 0.310, 0.595, 0.155, 0.070},  // Bt601_525
    {0.64, 0.33, 0.29, 0.60, 0.15, 0.06},      // Bt601_625
    {0.630, 0.340, 0.310, 0.595, 0.155, 0.070},  // Smpte240
    {0.708, 0.292, 0.170, 0.797, 0.131, 0.046},  // Bt2020
    {0.708, 0.292, 0.170, 0.797, 0.131, 0.046},  // Bt2100
};
// white points, XYZ with Y = 1 (ASTM E308, 2 degree observer; palette::white_point)
const double WP[11][3] = {
    {1.09850, 1.0, 0.35585}, {0.99072, 1.0, 0.85223}, {0.98074, 1.0, 1.18232}, {0.96422, 1.0, 0.82521},
    {0.95682, 1.0, 0.92149}, {0.95047, 1.0, 1.08883}, {0.94972, 1.0, 1.22638}, {1.0, 1.0, 1.0},
    {0.99186, 1.0, 0.67393}, {0.95041, 1.0, 1.08747}, {1.00962, 1.0, 0.64350}};
const double CONE[3][9] = {
    {0.8951, 0.2664, -0.1614, -0.7502, 1.7135, 0.0367, 0.0389, -0.0685, 1.0296},  // Bradford
    {0.40024, 0.7076, -0.08081, -0.2263, 1.16532, 0.0457, 0.0, 0.0, 0.91822},      // VonKries (HPE)
    {1, 0, 0, 0, 1, 0, 0, 0, 1}};                                                   // XyzScaling

bool to_xyz_d(uint32_t prim, uint32_t wp, double* o) {
  if (prim > 5 || wp > 10) return false;
  const double* p = PRIM[prim];
  double xr = p[0], yr = p[1], xg = p[2], yg = p[3], xb = p[4], yb = p[5];
  double P[9] = {xr / yr, xg / yg, xb / yb, 1.0, 1.0, 1.0, (1 - xr - yr) / yr, (1 - xg - yg) / yg, (1 - xb - yb) / yb};
  double Pi[9];
  inv3(P, Pi);
  const double* w = WP[wp];
  double S[3];
  for (int r = 0; r < 3; r++) S[r] = Pi[3 * r] * w[0] + Pi[3 * r + 1] * w[1] + Pi[3 * r + 2] * w[2];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o[3 * r + c] = P[3 * r + c] * S[c];
  return true;
}
bool adaptation_d(uint32_t method, uint32_t src, uint32_t dst, double* o) {
  if (src > 10 || dst > 10) return false;
  int mi = method == ZOSH_ADAPT_BRADFORD_VONKRIES ? 0 : method == ZOSH_ADAPT_VONKRIES ? 1 : method == ZOSH_ADAPT_XYZ ? 2 : -1;
  if (mi < 0) return false;
  const double* Mc = CONE[mi];
  double cs[3], cd[3];
  for (int r = 0; r < 3; r++) {
    cs[r] = Mc[3 * r] * WP[src][0] + Mc[3 * r + 1] * WP[src][1] + Mc[3 * r + 2] * WP[src][2];
    cd[r] = Mc[3 * r] * WP[dst][0] + Mc[3 * r + 1] * WP[dst][1] + Mc[3 * r + 2] * WP[dst][2];
  }
  double D[9] = {cd[0] / cs[0], 0, 0, 0, cd[1] / cs[1], 0, 0, 0, cd[2] / cs[2]};
  double Mi[9], t[9];
  inv3(Mc, Mi);
  mul3(D, Mc, t);
  mul3(Mi, t, o);
  return true;
}

// RowMatrix::multiply_right in f32: dot products left to right (color_matrix.rs:124-144)
void mul3_f32(const float* a, const float* b, float* o) {
  float t[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = (a[3 * r] * b[c] + a[3 * r + 1] * b[3 + c]) + a[3 * r + 2] * b[6 + c];
  memcpy(o, t, sizeof t);
}

bool same_chroma(const zos_desc& a, const zos_desc& b) {
  return a.block == b.block && a.bits == b.bits && a.parts == b.parts && a.color == b.color && a.transfer == b.transfer &&
         a.primaries == b.primaries && a.whitepoint == b.whitepoint;
}
void fix_layout(zos_desc& d) {
  d.texel_stride = d.block == ZOS_BLOCK_PIXEL ? zos_bits_bytes(d.bits) : 1;
  d.row_stride = zos_aligned_row_stride(d.width, d.texel_stride);
}
zos_step make_step(uint32_t kind, const double* m, const double* v = nullptr) {
  zos_step s;
  memset(&s, 0, sizeof s);
  s.kind = kind;
  if (m) to_f32(m, s.m);
  if (v) for (int i = 0; i < 3; i++) s.v[i] = (float)v[i];
  return s;
}

}  // namespace

// One recorded builder call of a TEMPLATE command buffer (one that declared a generic, command.rs:856-870): the
// descriptors of its registers depend on the types the caller binds, so its operations are kept as calls and built
// when `zosh_cb_invoke` replays them into the caller (monomorphisation by inlining; the reference monomorphises at
// link time, command.rs:2083-2185).  Register operands are positions in the record.
enum CallFn : uint32_t {
  FN_INPUT, FN_INPUT_GENERIC, FN_OUTPUT, FN_COLOR_CONVERT, FN_CHROMATIC_ADAPTATION, FN_INSCRIBE, FN_BLEND, FN_CROP, FN_AFFINE, FN_RESIZE,
  FN_TRANSMUTE, FN_BILINEAR, FN_SOLID_RGBA, FN_NORMAL2D, FN_FRACTAL_NOISE, FN_DERIVATIVE, FN_PALETTE, FN_EXTRACT, FN_INJECT, FN_BUFFER_INIT,
  FN_BUFFER_ZERO, FN_FROM_BUFFER, FN_WITH_BUFFER_BILINEAR, FN_DYNAMIC,
  FN_INVOKE,         // a call of another function from inside a template: i = function variable, generics / args below
  FN_INVOKED_RESULT  // Op::InvokedResult (command.rs:2821-2869): result u[0] of the FN_INVOKE at record position r[0]
};
struct Call {
  uint32_t fn = 0;
  int32_t r[2] = {-1, -1};  // register operands (record positions), -1 = none
  zos_desc d{};             // the descriptor argument, if the builder takes one
  uint32_t u[3] = {0, 0, 0};
  int32_t i = 0;
  float f[24] = {};
  zosh_rect rect{};
  std::vector<uint8_t> blob;  // buffer_init bytes / params of a user operator
  uint64_t len = 0;
  std::string source;
  bool knob = false;          // with_knob() preceded the call
  std::vector<zos_desc> generics;  // FN_INVOKE: a descriptor, or ZOSH_GENERIC_VAR | k in `reserved` = the template's generic k
  std::vector<int32_t> args;       // FN_INVOKE: argument registers (record positions)
};
struct zosh_signature {  // command::CommandSignature of a template, with the callee travelling inside
  std::shared_ptr<const std::vector<Call>> record;
  uint32_t num_generics = 0, num_inputs = 0, num_outputs = 0;
  const zosh_cb* origin = nullptr;  // identity of the template, checked by zosh_link
  std::shared_ptr<const std::vector<zosh_signature>> functions;  // the functions the template itself declared (nested calls)
};
struct KnobNote {  // a knob handed out while a function was inlined: which template, which of its registers, which id
  const zosh_cb* origin;
  int32_t pos;
  uint32_t knob;
};
struct zosh_cb {
  std::vector<zos_op> ops;  // op i defines register i (outputs define a register too, like the reference)
  uint32_t next_knob = 0, pending_knob = 0;
  std::vector<KnobNote> knob_notes;  // knobs of inlined calls, addressed as RegisterKnob{link_idx, register} after linking (command.rs:2134-2145)
  std::vector<std::shared_ptr<std::vector<uint8_t>>> blobs;  // initial bytes of buffer registers (zos_op::data points into them)
  bool is_template = false;
  uint32_t num_generics = 0;
  std::vector<Call> record;               // template only
  std::vector<zosh_signature> functions;  // FunctionVar i = functions[i] (command.rs:907-922)
};
struct KnobEntry { uint32_t link_idx; int32_t reg; uint32_t knob; };
struct zosh_program {
  std::vector<KnobEntry> knobs;  // RegisterKnob{link_idx, register} -> Knob (command.rs:701-705, 2134-2145); a later instantiation wins
  std::vector<zos_op> ops;
  std::vector<std::shared_ptr<std::vector<uint8_t>>> blobs;
  std::vector<int32_t> regmap;  // generic entry point: register of the template `main` -> register of its monomorphic copy
};

namespace {
// with_knob() marks the NEXT operation (command.rs:1865-1874).  A builder that fails must not leave the mark behind for an unrelated
// later call: every builder holds one of these, and a return without a pushed operation / recorded call clears the pending knob.
struct KnobGuard {
  zosh_cb* cb;
  size_t before;
  explicit KnobGuard(zosh_cb* c) : cb(c), before(c ? c->ops.size() + c->record.size() : 0) {}
  ~KnobGuard() { if (cb && cb->ops.size() + cb->record.size() == before) cb->pending_knob = 0; }
};
// an IMAGE register (outputs and byte buffers are not: the reference answers TYPE_ERR / BAD_REGISTER for them)
bool valid_reg(const zosh_cb* cb, int32_t r) { return r >= 0 && (size_t)r < cb->ops.size() && cb->ops[r].kind != ZOS_OP_OUTPUT && cb->ops[r].kind != ZOS_OP_BUFFER_INIT; }
bool buffer_reg(const zosh_cb* cb, int32_t r) { return r >= 0 && (size_t)r < cb->ops.size() && cb->ops[r].kind == ZOS_OP_BUFFER_INIT; }
zos_op new_op(zosh_cb* cb, uint32_t kind, int32_t s0, int32_t s1, const zos_desc& d) {
  zos_op op;
  memset(&op, 0, sizeof op);
  op.kind = kind;
  op.src[0] = s0; op.src[1] = s1;
  op.dst = (int32_t)cb->ops.size();
  op.reg = op.dst;
  op.desc = d;
  op.knob = cb->pending_knob;
  cb->pending_knob = 0;
  return op;
}
int32_t push(zosh_cb* cb, const zos_op& op, int32_t* reg) {
  cb->ops.push_back(op);
  if (reg) *reg = op.dst;
  return ZOSH_OK;
}
// ---- template command buffers: builders record instead of building ----
bool recording(const zosh_cb* cb) { return cb && cb->is_template; }
Call call(uint32_t fn, int32_t r0 = -1, int32_t r1 = -1) {
  Call c;
  c.fn = fn; c.r[0] = r0; c.r[1] = r1;
  return c;
}
Call call_d(uint32_t fn, const zos_desc* d, int32_t r0 = -1, int32_t r1 = -1) {
  Call c = call(fn, r0, r1);
  if (d) c.d = *d;
  return c;
}
Call call_f(Call c, const float* f, int n, int at = 0) {
  if (f) memcpy(c.f + at, f, sizeof(float) * n);
  return c;
}
Call call_u(Call c, uint32_t u0, uint32_t u1 = 0, uint32_t u2 = 0) {
  c.u[0] = u0; c.u[1] = u1; c.u[2] = u2;
  return c;
}
int32_t record(zosh_cb* cb, Call c, int32_t* reg) {
  for (int k = 0; k < 2; k++)  // operands name earlier, non-output entries of the record
    if (c.r[k] != -1 && (c.r[k] < 0 || (size_t)c.r[k] >= cb->record.size() || cb->record[c.r[k]].fn == FN_OUTPUT || cb->record[c.r[k]].fn == FN_INVOKE))
      return err(ZOSH_ERR_OTHER, "bad register");
  c.knob = cb->pending_knob != 0;
  cb->pending_knob = 0;
  cb->record.push_back(std::move(c));
  if (reg) *reg = (int32_t)cb->record.size() - 1;
  return ZOSH_OK;
}
}  // namespace

extern "C" {

const char* zosh_last_error(void) { return g_err.c_str(); }

int32_t zosh_to_xyz_matrix(uint32_t primaries, uint32_t whitepoint, float out[9]) {
  double m[9];
  if (!out) return err(ZOSH_ERR_OTHER, "null argument");
  if (!to_xyz_d(primaries, whitepoint, m)) return err(ZOSH_ERR_OTHER, "unknown primaries / whitepoint");
  to_f32(m, out);
  return ZOSH_OK;
}
int32_t zosh_adaptation_matrix(uint32_t method, uint32_t s, uint32_t d, float out[9]) {
  double m[9];
  if (!out) return err(ZOSH_ERR_OTHER, "null argument");
  if (method == ZOSH_ADAPT_BRADFORD_NONLINEAR) return err(ZOSH_ERR_UNIMPLEMENTED, "BradfordNonLinear (command.rs:3327-3331)");
  if (!adaptation_d(method, s, d, m)) return err(ZOSH_ERR_OTHER, "unknown adaptation method / whitepoint");
  to_f32(m, out);
  return ZOSH_OK;
}
int32_t zosh_whitepoint_xyz(uint32_t wp, float out[3]) {
  if (!out) return err(ZOSH_ERR_OTHER, "null argument");
  if (wp > 10) return err(ZOSH_ERR_OTHER, "unknown whitepoint");
  for (int i = 0; i < 3; i++) out[i] = (float)WP[wp][i];
  return ZOSH_OK;
}
void zosh_affine_identity(float m[9]) { if (!m) return; const float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; memcpy(m, id, sizeof id); }
void zosh_affine_scale(float m[9], float x, float y) { if (!m) return; const float p[9] = {x, 0, 0, 0, y, 0, 0, 0, 1}; mul3_f32(p, m, m); }
void zosh_affine_rotate(float m[9], float rad) {
  if (!m) return;
  const float c = cosf(rad), s = sinf(rad);
  const float p[9] = {c, s, 0, -s, c, 0, 0, 0, 1};
  mul3_f32(p, m, m);
}
void zosh_affine_shift(float m[9], float x, float y) { if (!m) return; const float p[9] = {1, 0, x, 0, 1, y, 0, 0, 1}; mul3_f32(p, m, m); }
zosh_rect zosh_rect_normalize(zosh_rect r) {
  uint32_t w = r.max_x > r.x ? r.max_x - r.x : 0;
  return zosh_rect{r.x, r.y, r.x + w, r.y + w};  // sic: max_y = y + width(), command.rs:3536-3543
}

zosh_cb* zosh_cb_new(void) { return new zosh_cb(); }
void zosh_cb_free(zosh_cb* cb) { delete cb; }

int32_t zosh_cb_with_knob(zosh_cb* cb) {
  if (!cb) return 0;
  cb->pending_knob = ++cb->next_knob;
  return (int32_t)cb->pending_knob;
}

uint32_t zosh_cb_num_ops(const zosh_cb* cb) { return cb ? (uint32_t)cb->ops.size() : 0; }
int32_t zosh_cb_describe(const zosh_cb* cb, int32_t reg, zos_desc* out) {
  if (!cb || !out || !valid_reg(cb, reg)) return err(ZOSH_ERR_OTHER, "bad register");
  *out = cb->ops[reg].desc;
  return ZOSH_OK;
}

int32_t zosh_cb_input(zosh_cb* cb, const zos_desc* desc, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc) return record(cb, call_d(FN_INPUT, desc), reg);
  if (!cb || !desc) return err(ZOSH_ERR_OTHER, "null argument");
  zos_desc d = *desc;
  if (d.block == ZOS_BLOCK_PIXEL && d.texel_stride != zos_bits_bytes(d.bits))
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "inconsistent input declared");  // command.rs:744-748
  if (d.width == 0 || d.height == 0) return err(ZOSH_ERR_BAD_DESCRIPTOR, "empty input declared");
  fix_layout(d);
  return push(cb, new_op(cb, ZOS_OP_INPUT, -1, -1, d), reg);
}

int32_t zosh_cb_output(zosh_cb* cb, int32_t src, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) return record(cb, call(FN_OUTPUT, src), reg);
  if (!cb || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  zos_op op = new_op(cb, ZOS_OP_OUTPUT, src, -1, cb->ops[src].desc);
  if (reg) *reg = op.dst;
  op.dst = -1;
  cb->ops.push_back(op);
  return ZOSH_OK;
}

int32_t zosh_cb_color_convert(zosh_cb* cb, int32_t src, const zos_desc* target, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && target) return record(cb, call_d(FN_COLOR_CONVERT, target, src), reg);
  if (!cb || !target || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& s = cb->ops[src].desc;
  zos_desc d = *target;
  d.width = s.width; d.height = s.height; d.block = ZOS_BLOCK_PIXEL;
  fix_layout(d);
  double T[9], Ti[9];
  zos_step st;
  if (s.color == ZOS_COLOR_RGB && d.color == ZOS_COLOR_RGB && s.whitepoint == d.whitepoint) {
    // command.rs:1022-1025 names the matrices the wrong way round and to_shader (3223-3228) then
    // computes to_xyz(dst) * inv(to_xyz(src)); reproduced as is (identity when primaries match)
    double A[9], B[9], M[9];
    if (!to_xyz_d(d.primaries, d.whitepoint, A) || !to_xyz_d(s.primaries, s.whitepoint, B)) return err(ZOSH_ERR_OTHER, "unknown primaries");
    inv3(B, Ti);
    mul3(A, Ti, M);
    st = make_step(ZOS_STEP_MATRIX, M);
  } else if (s.color == ZOS_COLOR_RGB && d.color == ZOS_COLOR_OKLAB && s.whitepoint == ZOS_WP_D65) {
    to_xyz_d(s.primaries, ZOS_WP_D65, T);
    st = make_step(ZOS_STEP_OKLAB_ENC, T);
  } else if (s.color == ZOS_COLOR_OKLAB && d.color == ZOS_COLOR_RGB && d.whitepoint == ZOS_WP_D65) {
    to_xyz_d(d.primaries, ZOS_WP_D65, T);
    inv3(T, Ti);
    st = make_step(ZOS_STEP_OKLAB_DEC, Ti);
  } else if (s.color == ZOS_COLOR_RGB && d.color == ZOS_COLOR_SRLAB2) {
    to_xyz_d(s.primaries, s.whitepoint, T);
    st = make_step(ZOS_STEP_SRLAB2_ENC, T);
  } else if (s.color == ZOS_COLOR_SRLAB2 && d.color == ZOS_COLOR_RGB) {
    to_xyz_d(d.primaries, d.whitepoint, T);
    inv3(T, Ti);
    if (s.whitepoint > 10) return err(ZOSH_ERR_OTHER, "unknown whitepoint");
    st = make_step(ZOS_STEP_SRLAB2_DEC, Ti, WP[s.whitepoint]);
  } else {
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "No conversion");  // command.rs:1077-1084
  }
  zos_op op = new_op(cb, ZOS_OP_PIXEL, src, -1, d);
  op.nsteps = 1;
  op.steps[0] = st;
  return push(cb, op, reg);
}

int32_t zosh_cb_chromatic_adaptation(zosh_cb* cb, int32_t src, uint32_t method, uint32_t target, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) return record(cb, call_u(call(FN_CHROMATIC_ADAPTATION, src), method, target), reg);
  if (!cb || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& s = cb->ops[src].desc;
  if (s.color != ZOS_COLOR_RGB) return err(ZOSH_ERR_BAD_DESCRIPTOR, "non-rgb chromatic adaptation");  // command.rs:1149-1157
  if (method == ZOSH_ADAPT_BRADFORD_NONLINEAR) return err(ZOSH_ERR_UNIMPLEMENTED, "BradfordNonLinear");
  double to[9], from[9], fi[9], ad[9], t[9], M[9];
  if (!to_xyz_d(s.primaries, s.whitepoint, to) || !to_xyz_d(s.primaries, target, from) || !adaptation_d(method, s.whitepoint, target, ad))
    return err(ZOSH_ERR_UNIMPLEMENTED, "whitepoint / method");
  inv3(from, fi);
  mul3(ad, to, t);
  mul3(fi, t, M);  // from_xyz(target) * adapt * to_xyz(source), command.rs:2527-2531
  zos_desc d = s;
  d.whitepoint = target;
  zos_op op = new_op(cb, ZOS_OP_PIXEL, src, -1, d);
  op.nsteps = 1;
  op.steps[0] = make_step(ZOS_STEP_MATRIX, M);
  return push(cb, op, reg);
}

static void compose_defaults(zos_compose_params& p) {
  memset(&p, 0, sizeof p);
  p.map = ZOS_MAP_RECT; p.sampling = ZOS_SAMPLE_NEAREST; p.blend = ZOS_BLEND_OVERWRITE; p.use_tma = 1;
}

int32_t zosh_cb_inscribe(zosh_cb* cb, int32_t below, zosh_rect rect, int32_t above, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) { Call c = call(FN_INSCRIBE, below, above); c.rect = rect; return record(cb, c, reg); }
  if (!cb || !valid_reg(cb, below) || !valid_reg(cb, above)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& b = cb->ops[below].desc;
  const zos_desc& a = cb->ops[above].desc;
  if (!same_chroma(a, b)) return err(ZOSH_ERR_CONFLICTING_TYPES, "inscribe: texel / colour of the layers differ");  // :1186-1190
  if (rect.x != 0 || rect.y != 0 || rect.max_x != a.width || rect.max_y != a.height) return err(ZOSH_ERR_OTHER, "inscribe: rect must be the layout of `above`");  // :1196-1198
  if (rect.max_x > b.width || rect.max_y > b.height) return err(ZOSH_ERR_OTHER, "inscribe: not contained in `below`");  // :1202-1206
  zosh_rect pl = zosh_rect_normalize(rect);
  zos_op op = new_op(cb, ZOS_OP_COMPOSE, below, above, b);
  compose_defaults(op.compose);
  op.compose.sel[2] = (int32_t)a.width; op.compose.sel[3] = (int32_t)a.height;
  op.compose.tgt[0] = (int32_t)pl.x; op.compose.tgt[1] = (int32_t)pl.y;
  op.compose.tgt[2] = (int32_t)(pl.max_x - pl.x); op.compose.tgt[3] = (int32_t)(pl.max_y - pl.y);
  return push(cb, op, reg);
}

int32_t zosh_cb_blend(zosh_cb* cb, int32_t below, zosh_rect rect, int32_t above, int32_t mode, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) { Call c = call(FN_BLEND, below, above); c.rect = rect; c.i = mode; return record(cb, c, reg); }
  // The reference returns UNIMPLEMENTED here (command.rs:1510-1519); semantics: DESIGN.md section 3.
  if (!cb || !valid_reg(cb, below) || !valid_reg(cb, above)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& b = cb->ops[below].desc;
  const zos_desc& a = cb->ops[above].desc;
  if (!same_chroma(a, b)) return err(ZOSH_ERR_CONFLICTING_TYPES, "blend: texel / colour of the layers differ");
  if (mode < ZOS_BLEND_CLEAR || mode > ZOS_BLEND_XOR) return err(ZOSH_ERR_OTHER, "blend: unknown mode");
  if (rect.max_x < rect.x || rect.max_y < rect.y || rect.max_x - rect.x != a.width || rect.max_y - rect.y != a.height)
    return err(ZOSH_ERR_OTHER, "blend: rect must have the size of `above`");
  if (rect.max_x > b.width || rect.max_y > b.height) return err(ZOSH_ERR_OTHER, "blend: not contained in `below`");
  if (a.color != ZOS_COLOR_RGB && a.color != ZOS_COLOR_SCALARS) return err(ZOSH_ERR_BAD_DESCRIPTOR, "blend: needs an RGB-ish colour");
  zos_op op = new_op(cb, ZOS_OP_COMPOSE, below, above, b);
  compose_defaults(op.compose);
  op.compose.blend = mode;
  op.compose.sel[2] = (int32_t)a.width; op.compose.sel[3] = (int32_t)a.height;
  op.compose.tgt[0] = (int32_t)rect.x; op.compose.tgt[1] = (int32_t)rect.y;
  op.compose.tgt[2] = (int32_t)a.width; op.compose.tgt[3] = (int32_t)a.height;
  return push(cb, op, reg);
}

int32_t zosh_cb_crop(zosh_cb* cb, int32_t src, zosh_rect rect, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) { Call c = call(FN_CROP, src); c.rect = rect; return record(cb, c, reg); }
  if (!cb || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& s = cb->ops[src].desc;
  if (rect.max_x <= rect.x || rect.max_y <= rect.y) return err(ZOSH_ERR_OTHER, "crop: empty rectangle");
  // the output keeps the source descriptor and the selection is stretched over it (command.rs:971-978, 2507-2526)
  zos_op op = new_op(cb, ZOS_OP_COMPOSE, -1, src, s);
  compose_defaults(op.compose);
  op.compose.sel[0] = (int32_t)rect.x; op.compose.sel[1] = (int32_t)rect.y;
  op.compose.sel[2] = (int32_t)(rect.max_x - rect.x); op.compose.sel[3] = (int32_t)(rect.max_y - rect.y);
  op.compose.tgt[2] = (int32_t)s.width; op.compose.tgt[3] = (int32_t)s.height;
  return push(cb, op, reg);
}

int32_t zosh_cb_affine(zosh_cb* cb, int32_t below, const float m[9], uint32_t sampling, int32_t above, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && m) return record(cb, call_u(call_f(call(FN_AFFINE, below, above), m, 9), sampling), reg);
  if (!cb || !m || !valid_reg(cb, below) || !valid_reg(cb, above)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& b = cb->ops[below].desc;
  const zos_desc& a = cb->ops[above].desc;
  if (!same_chroma(a, b)) return err(ZOSH_ERR_TYPE, "affine: texel / colour of the layers differ");  // command.rs:1646-1648
  // RowMatrix::det in f32 (color_matrix.rs:31-39), rejected below f32::EPSILON (command.rs:1650-1657)
  float det = m[0] * (m[4] * m[8] - m[7] * m[5]) - m[3] * (m[1] * m[8] - m[7] * m[2]) + m[6] * (m[1] * m[5] - m[4] * m[2]);
  if (!(fabsf(det) >= 1.1920929e-07f)) return err(ZOSH_ERR_OTHER, "affine: singular transformation");
  if (sampling > ZOS_SAMPLE_BILINEAR) return err(ZOSH_ERR_OTHER, "affine: unknown sampling");
  // AffineSample::BiLinear is UNIMPLEMENTED in the reference (command.rs:1659-1665); implemented here.
  if (sampling == ZOS_SAMPLE_BILINEAR && a.color != ZOS_COLOR_RGB && a.color != ZOS_COLOR_SCALARS)
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "affine: bilinear sampling needs an RGB-ish colour");
  double md[9], inv[9];
  for (int i = 0; i < 9; i++) md[i] = (double)m[i];
  inv3(md, inv);
  zos_op op = new_op(cb, ZOS_OP_COMPOSE, below, above, b);
  compose_defaults(op.compose);
  op.compose.map = ZOS_MAP_AFFINE;
  op.compose.sampling = (int32_t)sampling;
  to_f32(inv, op.compose.inv);
  return push(cb, op, reg);
}

int32_t zosh_cb_resize(zosh_cb* cb, int32_t below, uint32_t w, uint32_t h, uint32_t mode, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) return record(cb, call_u(call(FN_RESIZE, below), w, h, mode), reg);
  if (!cb || !valid_reg(cb, below)) return err(ZOSH_ERR_OTHER, "bad register");
  if (w == 0 || h == 0 || mode > ZOSH_RESIZE_BILINEAR) return err(ZOSH_ERR_OTHER, "resize: bad size / mode");
  zos_desc d = cb->ops[below].desc;
  if (d.block != ZOS_BLOCK_PIXEL) {  // a planar source resizes into its R'G'B' interpretation: caller converts first
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "resize: planar source, convert it first");
  }
  d.width = w; d.height = h;
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_COMPOSE, -1, below, d);
  compose_defaults(op.compose);
  op.compose.map = mode == ZOSH_RESIZE_REFERENCE ? ZOS_MAP_GRID8 : ZOS_MAP_SCALE;
  op.compose.sampling = mode == ZOSH_RESIZE_BILINEAR ? ZOS_SAMPLE_BILINEAR : ZOS_SAMPLE_NEAREST;
  return push(cb, op, reg);
}

int32_t zosh_cb_transmute(zosh_cb* cb, int32_t src, const zos_desc* target, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && target) return record(cb, call_d(FN_TRANSMUTE, target, src), reg);
  if (!cb || !target || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& s = cb->ops[src].desc;
  zos_desc d = *target;
  if (d.width != s.width || d.height != s.height) return err(ZOSH_ERR_BAD_DESCRIPTOR, "invalid transmute with mismatched size");  // :1305-1313
  if (d.block != ZOS_BLOCK_PIXEL || s.block != ZOS_BLOCK_PIXEL || zos_bits_bytes(d.bits) != zos_bits_bytes(s.bits))
    return err(ZOSH_ERR_CONFLICTING_TYPES, "transmute between texels of different size");  // :1329-1336
  if (d.texel_stride != zos_bits_bytes(d.bits)) return err(ZOSH_ERR_BAD_DESCRIPTOR, "invalid transmute with inconsistent result");
  fix_layout(d);
  return push(cb, new_op(cb, ZOS_OP_COPY, src, -1, d), reg);
}

int32_t zosh_cb_bilinear(zosh_cb* cb, const zos_desc* desc, const float p[24], int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc && p) return record(cb, call_f(call_d(FN_BILINEAR, desc), p, 24), reg);
  if (!cb || !desc || !p) return err(ZOSH_ERR_OTHER, "null argument");
  zos_desc d = *desc;
  if (d.block != ZOS_BLOCK_PIXEL || d.texel_stride != zos_bits_bytes(d.bits) || d.width == 0 || d.height == 0)
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "inconsistent descriptor for bilinear");
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_GENERATE, -1, -1, d);
  memcpy(op.gen, p, sizeof op.gen);
  return push(cb, op, reg);
}

int32_t zosh_cb_solid_rgba(zosh_cb* cb, const zos_desc* desc, const float color[4], int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc && color) return record(cb, call_f(call_d(FN_SOLID_RGBA, desc), color, 4), reg);
  if (!cb || !desc || !color) return err(ZOSH_ERR_OTHER, "null argument");
  zos_desc d = *desc;
  if (d.block != ZOS_BLOCK_PIXEL || d.texel_stride != zos_bits_bytes(d.bits) || d.width == 0 || d.height == 0)
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "inconsistent constant color image created");
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_GENERATE, -1, -1, d);
  memcpy(op.gen, color, 16);
  op.compose.map = 1;  // solid: the colour is written as is (solid_rgb.frag), not through mix()
  return push(cb, op, reg);
}

// shaders/distribution_normal2d.rs:25-100 and shaders/fractal_noise.rs:21-49: the parameter constructors, in f32 like the reference
void zosh_normal2d_with_diagonal(float var0, float var1, float out[7]) {
  if (!out) return;
  const float pi = 3.14159265358979323846f;
  const float d0 = var0 == 0.0f ? 0.0f : 1.0f / var0, d1 = var1 == 0.0f ? 0.0f : 1.0f / var1;
  const float f0 = var0 == 0.0f ? 1.0f : 2.0f * pi * var0, f1 = var1 == 0.0f ? 1.0f : 2.0f * pi * var1;
  out[0] = out[1] = 0.0f;
  out[2] = d0; out[3] = 0.0f; out[4] = 0.0f; out[5] = d1;
  out[6] = f0 * f1;
}
void zosh_normal2d_with_direction(float x, float y, float out[7]) {
  if (!out) return;
  auto sym = [](float a, float b) { float h = hypotf(a, b); float up = (1.0f / h) * (a / h); float low = a + b * (b / a); return up / low; };
  auto asym = [](float a, float b) { float lo = fminf(a, b), hi = fmaxf(a, b); float h = hypotf(lo, hi); float inner = fmaf(lo, lo / hi, hi); return ((lo / h) / inner) / h; };
  out[0] = out[1] = 0.0f;
  out[2] = sym(x, x); out[3] = asym(x, y); out[4] = asym(y, x); out[5] = sym(y, y);
  const float length_sq = (float)((double)x * (double)x + (double)y * (double)y);
  out[6] = 2.0f * 3.14159265358979323846f * length_sq;  // pseudo_determinant: 2.0 * PIf32 * length_sq, shaders/distribution_normal2d.rs:96
}
void zosh_fractal_noise_with_octaves(uint32_t octaves, float out[5]) {
  if (!out) return;
  out[0] = out[1] = 100.0f;
  out[2] = (float)(1.0 / (double)octaves);
  out[3] = 1.0f;
  out[4] = (float)octaves;
}
void zosh_fractal_noise_set_damping(float params[5], float damping) {
  if (!params) return;
  const float n = params[4];
  const float total = 1.0f - powf(damping, n);
  params[2] = fabsf(total) < 1e-7f ? 1.0f : (1.0f - damping) / total;
  params[3] = damping;
}

static int32_t push_generator(zosh_cb* cb, const zos_desc* desc, uint32_t kind, const float* p, int n, int32_t* reg, const char* what) {
  KnobGuard knob_guard(cb);
  if (!cb || !desc || !p) return err(ZOSH_ERR_OTHER, "null argument");
  zos_desc d = *desc;
  if (d.block != ZOS_BLOCK_PIXEL || d.texel_stride != zos_bits_bytes(d.bits) || d.width == 0 || d.height == 0) return err(ZOSH_ERR_BAD_DESCRIPTOR, what);
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_GENERATE, -1, -1, d);
  memcpy(op.gen, p, sizeof(float) * n);
  op.compose.map = (int32_t)kind;
  return push(cb, op, reg);
}
int32_t zosh_cb_distribution_normal2d(zosh_cb* cb, const zos_desc* desc, const float params[7], int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc && params) return record(cb, call_f(call_d(FN_NORMAL2D, desc), params, 7), reg);
  return push_generator(cb, desc, ZOS_GEN_NORMAL2D, params, 7, reg, "inconsistent descriptor for distribution_normal2d");
}
int32_t zosh_cb_distribution_fractal_noise(zosh_cb* cb, const zos_desc* desc, const float params[5], int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc && params) return record(cb, call_f(call_d(FN_FRACTAL_NOISE, desc), params, 5), reg);
  if (params && !(params[4] >= 0.0f && params[4] <= 64.0f)) return err(ZOSH_ERR_OTHER, "fractal noise: 0..64 octaves");
  return push_generator(cb, desc, ZOS_GEN_FRACTAL_NOISE, params, 5, reg, "inconsistent descriptor for distribution_fractal_noise");
}

// ---- byte buffers (command.rs:1777-1803, 937-968, 1963-2060; tests/buffer.rs)
static int32_t push_buffer(zosh_cb* cb, const void* data, uint64_t len, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (!cb) return err(ZOSH_ERR_OTHER, "null argument");
  if (len == 0 || len > (1ull << 32)) return err(ZOSH_ERR_OTHER, "buffer size out of range");
  zos_desc none;
  memset(&none, 0, sizeof none);
  zos_op op = new_op(cb, ZOS_OP_BUFFER_INIT, -1, -1, none);
  op.data_len = len;
  if (data) {
    auto blob = std::make_shared<std::vector<uint8_t>>((const uint8_t*)data, (const uint8_t*)data + len);
    cb->blobs.push_back(blob);
    op.data = blob->data();
  }
  return push(cb, op, reg);
}
int32_t zosh_cb_buffer_init(zosh_cb* cb, const void* data, uint64_t len, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && data && len && len <= (1ull << 32)) {
    Call c = call(FN_BUFFER_INIT);
    c.blob.assign((const uint8_t*)data, (const uint8_t*)data + len);
    return record(cb, std::move(c), reg);
  }
  if (!data) return err(ZOSH_ERR_OTHER, "null data");
  return push_buffer(cb, data, len, reg);
}
int32_t zosh_cb_buffer_zero(zosh_cb* cb, uint64_t len, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && len && len <= (1ull << 32)) { Call c = call(FN_BUFFER_ZERO); c.len = len; return record(cb, c, reg); }
  return push_buffer(cb, nullptr, len, reg);
}
int32_t zosh_cb_buffer_size(const zosh_cb* cb, int32_t reg, uint64_t* out) {
  if (recording(cb) && out && reg >= 0 && (size_t)reg < cb->record.size() &&
      (cb->record[reg].fn == FN_BUFFER_INIT || cb->record[reg].fn == FN_BUFFER_ZERO)) {  // a buffer's size does not depend on the bound types
    *out = cb->record[reg].fn == FN_BUFFER_INIT ? cb->record[reg].blob.size() : cb->record[reg].len;
    return ZOSH_OK;
  }
  if (!cb || !out || !buffer_reg(cb, reg)) return err(ZOSH_ERR_TYPE, "not a buffer register");
  *out = cb->ops[reg].data_len;
  return ZOSH_OK;
}
int32_t zosh_cb_from_buffer(zosh_cb* cb, int32_t buffer, const zos_desc* desc, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc) return record(cb, call_d(FN_FROM_BUFFER, desc, buffer), reg);
  if (!cb || !desc) return err(ZOSH_ERR_OTHER, "null argument");
  if (!buffer_reg(cb, buffer)) return err(ZOSH_ERR_TYPE, "from_buffer: not a buffer register (CommandError::TYPE_ERR)");
  zos_desc d = *desc;
  if (d.block != ZOS_BLOCK_PIXEL || d.texel_stride != zos_bits_bytes(d.bits) || d.width == 0 || d.height == 0)
    return err(ZOSH_ERR_OTHER, "from_buffer: descriptor has no aligned layout (CommandError::INVALID_CALL)");
  fix_layout(d);  // Descriptor::to_aligned
  if (cb->ops[buffer].data_len < d.row_stride * d.height) return err(ZOSH_ERR_OTHER, "from_buffer: buffer smaller than the aligned image (CommandError::INVALID_CALL)");
  return push(cb, new_op(cb, ZOS_OP_FROM_BUFFER, buffer, -1, d), reg);
}
int32_t zosh_cb_with_buffer_bilinear(zosh_cb* cb, int32_t buffer, const zos_desc* desc, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && desc) return record(cb, call_d(FN_WITH_BUFFER_BILINEAR, desc, buffer), reg);
  if (!cb || !desc) return err(ZOSH_ERR_OTHER, "null argument");
  if (!buffer_reg(cb, buffer)) return err(ZOSH_ERR_TYPE, "with_buffer: not a buffer register");
  if (cb->ops[buffer].data_len < 96) return err(ZOSH_ERR_OTHER, "with_buffer: the bilinear parameter block needs 96 bytes");
  zos_desc d = *desc;
  if (d.block != ZOS_BLOCK_PIXEL || d.texel_stride != zos_bits_bytes(d.bits) || d.width == 0 || d.height == 0)
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "inconsistent descriptor for bilinear");
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_GENERATE, buffer, -1, d);
  op.compose.map = ZOS_GEN_BILINEAR;
  return push(cb, op, reg);
}

// ---- user operators (command.rs:2933-3060 construct_dynamic / unary_dynamic / binary_dynamic; command/dynamic.rs)
int32_t zosh_cb_dynamic(zosh_cb* cb, int32_t src0, int32_t src1, const char* cuda_source, const zos_desc* desc, const void* params,
                        uint64_t params_len, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && cuda_source && desc && !(src0 < 0 && src1 >= 0)) {
    Call c = call_d(FN_DYNAMIC, desc, src0 < 0 ? -1 : src0, src1 < 0 ? -1 : src1);
    c.source = cuda_source;
    if (params && params_len) c.blob.assign((const uint8_t*)params, (const uint8_t*)params + params_len);
    return record(cb, std::move(c), reg);
  }
  if (!cb || !cuda_source || !desc) return err(ZOSH_ERR_OTHER, "null argument");
  if (src0 < 0 && src1 >= 0) return err(ZOSH_ERR_OTHER, "binary_dynamic needs both operands");
  for (int32_t r : {src0, src1})
    if (r >= 0 && !valid_reg(cb, r)) return err(ZOSH_ERR_OTHER, "dynamic operand is not an image register (CommandError::INVALID_CALL)");
  if (params_len > 4096 || (params_len && !params)) return err(ZOSH_ERR_OTHER, "dynamic operator: 0..4096 bytes of data");
  zos_desc d = *desc;
  if (d.block != ZOS_BLOCK_PIXEL || d.texel_stride != zos_bits_bytes(d.bits) || d.width == 0 || d.height == 0)
    return err(ZOSH_ERR_BAD_DESCRIPTOR, "inconsistent descriptor returned by the shader command");
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_DYNAMIC, src0, src1, d);
  const size_t n = strlen(cuda_source);
  auto text = std::make_shared<std::vector<uint8_t>>((const uint8_t*)cuda_source, (const uint8_t*)cuda_source + n + 1);
  cb->blobs.push_back(text);
  op.source = (const char*)text->data();
  if (params_len) {
    auto blob = std::make_shared<std::vector<uint8_t>>((const uint8_t*)params, (const uint8_t*)params + params_len);
    cb->blobs.push_back(blob);
    op.data = blob->data();
    op.data_len = params_len;
  }
  return push(cb, op, reg);
}

int32_t zosh_cb_derivative(zosh_cb* cb, int32_t src, uint32_t method, uint32_t height_direction, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) return record(cb, call_u(call(FN_DERIVATIVE, src), method, height_direction), reg);
  if (!cb || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  float sm[3];
  switch (method) {  // command.rs:3343-3409
    case ZOSH_DERIV_PREWITT: sm[0] = sm[1] = sm[2] = 1.0f / 3.0f; break;
    case ZOSH_DERIV_SOBEL: sm[0] = 0.25f; sm[1] = 0.5f; sm[2] = 0.25f; break;
    case ZOSH_DERIV_SCHARR3: sm[0] = sm[2] = (float)(46.84 / 256.0); sm[1] = (float)(162.32 / 256.0); break;
    case ZOSH_DERIV_SCHARR3_TO_4BIT: sm[0] = sm[2] = 3.0f / 16.0f; sm[1] = 10.0f / 16.0f; break;
    case ZOSH_DERIV_SCHARR3_TO_8BIT: sm[0] = sm[2] = 47.0f / 256.0f; sm[1] = 162.0f / 256.0f; break;
    default: return err(ZOSH_ERR_UNIMPLEMENTED, "derivative method (CompileError::NotYetImplemented)");
  }
  const float dd[3] = {0.5f, 0.0f, -0.5f};
  zos_op op = new_op(cb, ZOS_OP_BOX3, src, -1, cb->ops[src].desc);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      float v = sm[r] * dd[c];  // RowMatrix::with_outer_product
      op.gen[height_direction ? 3 * c + r : 3 * r + c] = v;
    }
  return push(cb, op, reg);
}

int32_t zosh_cb_palette(zosh_cb* cb, int32_t palette, int32_t indices, const float xc[4], const float yc[4], int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb) && xc && yc) return record(cb, call_f(call_f(call(FN_PALETTE, palette, indices), xc, 4), yc, 4, 4), reg);
  if (!cb || !xc || !yc || !valid_reg(cb, palette) || !valid_reg(cb, indices)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& p = cb->ops[palette].desc;
  zos_desc d = cb->ops[indices].desc;  // layout of the indices, chroma of the palette (command.rs:1467-1471)
  d.bits = p.bits; d.parts = p.parts; d.color = p.color; d.transfer = p.transfer; d.primaries = p.primaries; d.whitepoint = p.whitepoint;
  d.block = p.block;
  fix_layout(d);
  zos_op op = new_op(cb, ZOS_OP_PALETTE, palette, indices, d);
  memcpy(op.compose.inv, xc, 16);
  memcpy(op.compose.inv + 4, yc, 16);
  return push(cb, op, reg);
}

// TexelExt::channel_texel (buffer.rs:57-62): same bit depth per channel, parts = the single channel
static bool channel_texel(const zos_desc& s, uint32_t channel, zos_desc& d) {
  d = s;
  switch (s.bits) {
    case ZOS_BITS_UINT8X4: case ZOS_BITS_UINT8X3: case ZOS_BITS_UINT8X2: d.bits = ZOS_BITS_UINT8; break;
    case ZOS_BITS_UINT16X4: case ZOS_BITS_UINT16X3: case ZOS_BITS_UINT16X2: d.bits = ZOS_BITS_UINT16; break;
    default: return false;
  }
  switch (channel) {
    case 0: d.parts = ZOS_PARTS_R; break;
    case 1: d.parts = ZOS_PARTS_G; break;
    case 2: d.parts = ZOS_PARTS_B; break;
    case 3: d.parts = ZOS_PARTS_A; break;
    default: return false;
  }
  fix_layout(d);
  return true;
}

int32_t zosh_cb_extract(zosh_cb* cb, int32_t src, uint32_t channel, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) return record(cb, call_u(call(FN_EXTRACT, src), channel), reg);
  if (!cb || !valid_reg(cb, src)) return err(ZOSH_ERR_OTHER, "bad register");
  zos_desc d;
  if (cb->ops[src].desc.block != ZOS_BLOCK_PIXEL || !channel_texel(cb->ops[src].desc, channel, d)) return err(ZOSH_ERR_OTHER, "extract: no such channel texel");  // :1232-1235
  if (channel > 2) return err(ZOSH_ERR_OTHER, "extract: channel position");  // ChannelPosition::new knows R, G, B only (buffer.rs:186-194)
  // a full copy; the channel is picked when the result is packed into the single-channel texel (command.rs:2578-2597)
  return push(cb, new_op(cb, ZOS_OP_PIXEL, src, -1, d), reg);
}

int32_t zosh_cb_inject(zosh_cb* cb, int32_t below, uint32_t channel, int32_t above, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (recording(cb)) return record(cb, call_u(call(FN_INJECT, below, above), channel), reg);
  if (!cb || !valid_reg(cb, below) || !valid_reg(cb, above)) return err(ZOSH_ERR_OTHER, "bad register");
  const zos_desc& b = cb->ops[below].desc;
  const zos_desc& a = cb->ops[above].desc;
  zos_desc expect;
  if (b.block != ZOS_BLOCK_PIXEL || !channel_texel(b, channel, expect)) return err(ZOSH_ERR_OTHER, "inject: no such channel texel");  // :1392-1394
  if (channel > 2) return err(ZOSH_ERR_OTHER, "inject: channel position");
  float color[4] = {0, 0, 0, 0};
  switch (a.parts) {  // TexelExt::channel_weight_vec4, buffer.rs:64-81
    case ZOS_PARTS_R: case ZOS_PARTS_LUMA: color[0] = 1; break;
    case ZOS_PARTS_G: color[1] = 1; break;
    case ZOS_PARTS_B: color[2] = 1; break;
    case ZOS_PARTS_A: color[3] = 1; break;
    default: return err(ZOSH_ERR_CONFLICTING_TYPES, "inject: `above` must be a single-channel image");  // :1396-1405, 1414-1416
  }
  // everything but the sample parts must match the expected channel texel (command.rs:1407-1427)
  if (a.bits != expect.bits || a.color != expect.color || a.transfer != expect.transfer || a.primaries != expect.primaries ||
      a.whitepoint != expect.whitepoint || a.width != b.width || a.height != b.height)
    return err(ZOSH_ERR_CONFLICTING_TYPES, "inject: `above` does not match the channel texel of `below`");
  zos_op op = new_op(cb, ZOS_OP_COMPOSE, below, above, b);
  compose_defaults(op.compose);
  op.compose.blend = ZOS_BLEND_INJECT;
  op.compose.sel[2] = (int32_t)a.width; op.compose.sel[3] = (int32_t)a.height;
  op.compose.tgt[2] = (int32_t)a.width; op.compose.tgt[3] = (int32_t)a.height;
  op.compose.inject_mix[channel] = 1.0f;
  memcpy(op.compose.inject_color, color, sizeof color);
  return push(cb, op, reg);
}

// ---- functions and generics (command.rs:856-922, 2821-2869; linking :2083-2185) ----
int32_t zosh_cb_generic(zosh_cb* cb, int32_t* var) {
  if (!cb) return err(ZOSH_ERR_OTHER, "null argument");
  if (!cb->is_template && !cb->ops.empty()) return err(ZOSH_ERR_OTHER, "generics must be declared before the first operation");
  cb->is_template = true;
  if (var) *var = (int32_t)cb->num_generics;
  cb->num_generics++;
  return ZOSH_OK;
}
int32_t zosh_cb_input_generic(zosh_cb* cb, int32_t var, int32_t* reg) {
  KnobGuard knob_guard(cb);
  if (!cb || !cb->is_template || var < 0 || (uint32_t)var >= cb->num_generics) return err(ZOSH_ERR_OTHER, "input_generic: unknown generic");
  Call c = call(FN_INPUT_GENERIC);
  c.i = var;
  return record(cb, c, reg);
}
int32_t zosh_cb_computed_signature(const zosh_cb* cb, zosh_signature** out) {
  if (!cb || !out) return err(ZOSH_ERR_OTHER, "null argument");
  if (!cb->is_template) return err(ZOSH_ERR_UNIMPLEMENTED, "signatures of non-generic command buffers are not supported");
  zosh_signature* sig = new zosh_signature();
  sig->record = std::make_shared<const std::vector<Call>>(cb->record);
  sig->num_generics = cb->num_generics;
  for (const Call& c : cb->record) {
    sig->num_inputs += c.fn == FN_INPUT || c.fn == FN_INPUT_GENERIC;
    sig->num_outputs += c.fn == FN_OUTPUT;
  }
  sig->origin = cb;
  sig->functions = std::make_shared<const std::vector<zosh_signature>>(cb->functions);
  *out = sig;
  return ZOSH_OK;
}
void zosh_signature_free(zosh_signature* sig) { delete sig; }
uint32_t zosh_signature_num_generics(const zosh_signature* sig) { return sig ? sig->num_generics : 0; }
uint32_t zosh_signature_num_inputs(const zosh_signature* sig) { return sig ? sig->num_inputs : 0; }
uint32_t zosh_signature_num_outputs(const zosh_signature* sig) { return sig ? sig->num_outputs : 0; }
int32_t zosh_cb_function(zosh_cb* cb, const zosh_signature* sig, int32_t* function) {
  if (!cb || !sig) return err(ZOSH_ERR_OTHER, "null argument");
  cb->functions.push_back(*sig);
  if (function) *function = (int32_t)cb->functions.size() - 1;
  return ZOSH_OK;
}
uint32_t zosh_cb_num_functions(const zosh_cb* cb) { return cb ? (uint32_t)cb->functions.size() : 0; }

static int32_t replay(zosh_cb* cb, const Call& c, int32_t r0, int32_t r1, int32_t* reg) {
  if (c.knob) zosh_cb_with_knob(cb);
  switch (c.fn) {
    case FN_COLOR_CONVERT: return zosh_cb_color_convert(cb, r0, &c.d, reg);
    case FN_CHROMATIC_ADAPTATION: return zosh_cb_chromatic_adaptation(cb, r0, c.u[0], c.u[1], reg);
    case FN_INSCRIBE: return zosh_cb_inscribe(cb, r0, c.rect, r1, reg);
    case FN_BLEND: return zosh_cb_blend(cb, r0, c.rect, r1, c.i, reg);
    case FN_CROP: return zosh_cb_crop(cb, r0, c.rect, reg);
    case FN_AFFINE: return zosh_cb_affine(cb, r0, c.f, c.u[0], r1, reg);
    case FN_RESIZE: return zosh_cb_resize(cb, r0, c.u[0], c.u[1], c.u[2], reg);
    case FN_TRANSMUTE: return zosh_cb_transmute(cb, r0, &c.d, reg);
    case FN_BILINEAR: return zosh_cb_bilinear(cb, &c.d, c.f, reg);
    case FN_SOLID_RGBA: return zosh_cb_solid_rgba(cb, &c.d, c.f, reg);
    case FN_NORMAL2D: return zosh_cb_distribution_normal2d(cb, &c.d, c.f, reg);
    case FN_FRACTAL_NOISE: return zosh_cb_distribution_fractal_noise(cb, &c.d, c.f, reg);
    case FN_DERIVATIVE: return zosh_cb_derivative(cb, r0, c.u[0], c.u[1], reg);
    case FN_PALETTE: return zosh_cb_palette(cb, r0, r1, c.f, c.f + 4, reg);
    case FN_EXTRACT: return zosh_cb_extract(cb, r0, c.u[0], reg);
    case FN_INJECT: return zosh_cb_inject(cb, r0, c.u[0], r1, reg);
    case FN_BUFFER_INIT: return zosh_cb_buffer_init(cb, c.blob.data(), c.blob.size(), reg);
    case FN_BUFFER_ZERO: return zosh_cb_buffer_zero(cb, c.len, reg);
    case FN_FROM_BUFFER: return zosh_cb_from_buffer(cb, r0, &c.d, reg);
    case FN_WITH_BUFFER_BILINEAR: return zosh_cb_with_buffer_bilinear(cb, r0, &c.d, reg);
    case FN_DYNAMIC: return zosh_cb_dynamic(cb, r0, r1, c.source.c_str(), &c.d, c.blob.empty() ? nullptr : c.blob.data(), c.blob.size(), reg);
    default: return err(ZOSH_ERR_OTHER, "invoke: unknown recorded call");
  }
}

// Inline `sig` into the (non-template) command buffer `cb`.  Type errors of the callee under the bound types surface here.
static int32_t inline_signature(zosh_cb* cb, const zosh_signature& sig, const zos_desc* generics, uint32_t num_generics,
                                const int32_t* arguments, uint32_t num_arguments, std::vector<int32_t>& results, int depth,
                                std::vector<int32_t>* entry_map = nullptr) {
  // entry_map != nullptr: `sig` is a generic ENTRY POINT -- its inputs and outputs become Input / Output operations of cb
  // (instead of the caller's registers), and the map from its registers to cb's is handed back.
  if (entry_map) num_arguments = sig.num_inputs;
  if (depth > 32) return err(ZOSH_ERR_OTHER, "invoke: functions nest deeper than 32 calls (recursion?)");
  if (num_generics != sig.num_generics || num_arguments != sig.num_inputs)
    return err(ZOSH_ERR_TYPE, "invoke: number of generics / arguments differs from the signature (CommandError::INVALID_CALL)");
  const std::vector<Call>& rec = *sig.record;
  // arguments are type checked before anything is pushed
  uint32_t nxt = 0;
  for (const Call& c : rec) {
    if (entry_map || (c.fn != FN_INPUT && c.fn != FN_INPUT_GENERIC)) continue;
    const int32_t real = arguments[nxt++];
    if (!valid_reg(cb, real)) return err(ZOSH_ERR_OTHER, "invoke: bad argument register");
    zos_desc want = c.fn == FN_INPUT_GENERIC ? generics[c.i] : c.d;
    const zos_desc& have = cb->ops[real].desc;
    if (!same_chroma(have, want) || have.width != want.width || have.height != want.height)
      return err(ZOSH_ERR_TYPE, "invoke: an argument does not have the declared type (CommandError::INVALID_CALL)");
  }
  std::vector<int32_t> map(rec.size(), -1);
  std::vector<std::vector<int32_t>> nested(rec.size());  // results of the FN_INVOKE at each position
  nxt = 0;
  for (size_t pos = 0; pos < rec.size(); pos++) {
    const Call& c = rec[pos];
    if (c.fn == FN_INPUT || c.fn == FN_INPUT_GENERIC) {
      if (!entry_map) { map[pos] = arguments[nxt++]; continue; }
      if (c.knob) zosh_cb_with_knob(cb);
      const int32_t st = zosh_cb_input(cb, c.fn == FN_INPUT_GENERIC ? &generics[c.i] : &c.d, &map[pos]);
      if (st != ZOSH_OK) return st;
      if (c.knob) cb->knob_notes.push_back(KnobNote{sig.origin, (int32_t)pos, cb->next_knob});
      continue;
    }
    const int32_t r0 = c.r[0] >= 0 ? map[c.r[0]] : -1, r1 = c.r[1] >= 0 ? map[c.r[1]] : -1;
    if (c.fn == FN_OUTPUT) {
      results.push_back(r0);
      if (entry_map) {
        const int32_t st = zosh_cb_output(cb, r0, &map[pos]);
        if (st != ZOSH_OK) return st;
      }
      continue;
    }
    if (c.fn == FN_INVOKED_RESULT) { map[pos] = nested[c.r[0]][c.u[0]]; continue; }
    if (c.fn == FN_INVOKE) {
      if (c.i < 0 || (size_t)c.i >= sig.functions->size()) return err(ZOSH_ERR_OTHER, "invoke: unknown function");
      std::vector<zos_desc> gens;
      for (const zos_desc& g : c.generics)  // the callee's own generics flow into the nested call
        gens.push_back((g.reserved & ZOSH_GENERIC_VAR) ? generics[g.reserved & ~ZOSH_GENERIC_VAR] : g);
      std::vector<int32_t> args;
      for (int32_t a : c.args) args.push_back(map[a]);
      const int32_t st = inline_signature(cb, (*sig.functions)[c.i], gens.data(), (uint32_t)gens.size(), args.data(), (uint32_t)args.size(),
                                          nested[pos], depth + 1);
      if (st != ZOSH_OK) return st;
      continue;
    }
    const int32_t st = replay(cb, c, r0, r1, &map[pos]);
    if (st != ZOSH_OK) return st;
    if (c.knob) cb->knob_notes.push_back(KnobNote{sig.origin, (int32_t)pos, cb->next_knob});  // the id replay() just handed out
  }
  if (entry_map) *entry_map = map;
  return ZOSH_OK;
}

int32_t zosh_cb_invoke(zosh_cb* cb, int32_t function, const zos_desc* generics, uint32_t num_generics, const int32_t* arguments,
                       uint32_t num_arguments, int32_t* results, uint32_t results_cap, uint32_t* num_results) {
  KnobGuard knob_guard(cb);
  if (!cb || (num_generics && !generics) || (num_arguments && !arguments)) return err(ZOSH_ERR_OTHER, "null argument");
  if (function < 0 || (size_t)function >= cb->functions.size()) return err(ZOSH_ERR_OTHER, "invoke: unknown function");  // BAD_REGISTER
  const zosh_signature sig = cb->functions[function];
  if (results_cap < sig.num_outputs) return err(ZOSH_ERR_OTHER, "invoke: results array too small");
  if (cb->is_template) {  // recorded; types are checked when the enclosing template is itself inlined
    if (num_generics != sig.num_generics || num_arguments != sig.num_inputs)
      return err(ZOSH_ERR_TYPE, "invoke: number of generics / arguments differs from the signature (CommandError::INVALID_CALL)");
    Call c = call(FN_INVOKE);
    c.i = function;
    for (uint32_t k = 0; k < num_generics; k++) {
      if ((generics[k].reserved & ZOSH_GENERIC_VAR) && (generics[k].reserved & ~ZOSH_GENERIC_VAR) >= cb->num_generics)
        return err(ZOSH_ERR_OTHER, "invoke: unknown generic");
      c.generics.push_back(generics[k]);
    }
    for (uint32_t k = 0; k < num_arguments; k++) {
      const int32_t a = arguments[k];
      if (a < 0 || (size_t)a >= cb->record.size() || cb->record[a].fn == FN_OUTPUT || cb->record[a].fn == FN_INVOKE)
        return err(ZOSH_ERR_OTHER, "invoke: bad argument register");
      c.args.push_back(a);
    }
    int32_t at = -1;
    record(cb, std::move(c), &at);
    for (uint32_t k = 0; k < sig.num_outputs; k++) {
      Call r = call_u(call(FN_INVOKED_RESULT), k);
      r.r[0] = at;
      cb->record.push_back(r);
      results[k] = (int32_t)cb->record.size() - 1;
    }
    if (num_results) *num_results = sig.num_outputs;
    return ZOSH_OK;
  }
  const size_t ops_before = cb->ops.size(), blobs_before = cb->blobs.size(), notes_before = cb->knob_notes.size();
  const uint32_t knob_before = cb->next_knob;
  std::vector<int32_t> out;
  const int32_t st = inline_signature(cb, sig, generics, num_generics, arguments, num_arguments, out, 0);
  if (st != ZOSH_OK) {  // a failed call leaves the caller untouched: undo the partial inlining
    cb->ops.resize(ops_before);
    cb->blobs.resize(blobs_before);
    cb->knob_notes.resize(notes_before);
    cb->next_knob = knob_before;
    cb->pending_knob = 0;
    return st;
  }
  for (size_t k = 0; k < out.size(); k++) results[k] = out[k];
  if (num_results) *num_results = (uint32_t)out.size();
  return ZOSH_OK;
}

int32_t zosh_link(const zosh_cb* main_cb, const zos_desc* tys, uint32_t num_tys, const zosh_cb* const* functions, uint32_t num_functions,
                  const uint32_t* links, const uint32_t* links_per_program, zosh_program** out) {
  if (!main_cb || !out || (num_functions && !functions) || !links_per_program || (num_tys && !tys)) return err(ZOSH_ERR_OTHER, "null argument");
  if (num_tys != main_cb->num_generics)
    return err(ZOSH_ERR_TYPE, "link: one type per generic of the entry point (CommandError::TYPE_ERR)");
  const uint32_t* table = links;
  for (uint32_t p = 0; p <= num_functions; p++) {
    const zosh_cb* prog = p == 0 ? main_cb : functions[p - 1];
    if (!prog) return err(ZOSH_ERR_OTHER, "link: null program");
    if (links_per_program[p] != prog->functions.size()) return err(ZOSH_ERR_OTHER, "link: one link per declared function of every program");
    for (uint32_t f = 0; f < links_per_program[p]; f++) {
      const uint32_t target = table[f];
      if (target < 1 || target > num_functions) return err(ZOSH_ERR_OTHER, "link: bad function index");
      if (functions[target - 1] != prog->functions[f].origin)
        return err(ZOSH_ERR_TYPE, "link: the linked function has another signature (CommandError::TYPE_ERR)");
    }
    table += links_per_program[p];
  }
  // RegisterKnob{link_idx, register} -> Knob: link 0 = main's own operations, link k = the template functions[k - 1]
  auto knob_table = [&](const zosh_cb& built, bool main_is_template, zosh_program* prog) {
    if (!main_is_template)
      for (const zos_op& op : built.ops)
        if (op.knob) prog->knobs.push_back(KnobEntry{0, op.reg, op.knob});
    for (const KnobNote& n : built.knob_notes) {
      uint32_t link = n.origin == main_cb ? 0 : ~0u;
      for (uint32_t k = 0; k < num_functions && link == ~0u; k++)
        if (functions[k] == n.origin) link = k + 1;
      if (link != ~0u) prog->knobs.push_back(KnobEntry{link, n.pos, n.knob});
    }
  };
  if (!main_cb->is_template) {
    const int32_t st0 = zosh_compile(main_cb, out);
    if (st0 == ZOSH_OK) { (*out)->knobs.clear(); knob_table(*main_cb, false, *out); }
    return st0;
  }
  // generic entry point: build the monomorphic copy of `main` under `tys`, compile that, remember how registers map
  zosh_signature* sig = nullptr;
  int32_t st = zosh_cb_computed_signature(main_cb, &sig);
  if (st != ZOSH_OK) return st;
  zosh_cb mono;
  std::vector<int32_t> results, map;
  st = inline_signature(&mono, *sig, tys, num_tys, nullptr, 0, results, 0, &map);
  delete sig;
  if (st != ZOSH_OK) return st;
  st = zosh_compile(&mono, out);
  if (st == ZOSH_OK) {
    (*out)->regmap = map;
    (*out)->knobs.clear();
    knob_table(mono, true, *out);
  }
  return st;
}
// Executable::query_knob (run.rs:1016ff; RegisterKnob, command.rs:701-705): 0 = that register has no knob
uint32_t zosh_program_knob(const zosh_program* p, uint32_t link_idx, int32_t reg) {
  uint32_t found = 0;
  if (p)
    for (const KnobEntry& e : p->knobs)
      if (e.link_idx == link_idx && e.reg == reg) found = e.knob;
  return found;
}
int32_t zosh_program_register(const zosh_program* p, int32_t reg) {
  if (!p || reg < 0) return -1;
  if (p->regmap.empty()) return reg;
  return (size_t)reg < p->regmap.size() ? p->regmap[reg] : -1;
}

int32_t zosh_compile(const zosh_cb* cb, zosh_program** out) {
  if (!cb || !out) return err(ZOSH_ERR_OTHER, "null argument");
  if (cb->is_template) return err(ZOSH_ERR_UNIMPLEMENTED, "generic entry points are not supported (CommandError::UNIMPLEMENTED)");
  // liveness (command.rs:2216-2291): only operations that reach an output are emitted
  std::vector<char> live(cb->ops.size(), 0);
  for (size_t i = cb->ops.size(); i-- > 0;) {
    const zos_op& op = cb->ops[i];
    if (op.kind == ZOS_OP_OUTPUT || op.kind == ZOS_OP_INPUT) live[i] = 1;
    if (!live[i]) continue;
    for (int s = 0; s < 2; s++)
      if (op.src[s] >= 0) live[op.src[s]] = 1;
  }
  zosh_program* p = new zosh_program();
  p->blobs = cb->blobs;  // zos_op::data of buffer registers points into these
  for (size_t i = 0; i < cb->ops.size(); i++)
    if (live[i]) p->ops.push_back(cb->ops[i]);
  for (const zos_op& op : cb->ops)
    if (op.knob) p->knobs.push_back(KnobEntry{0, op.reg, op.knob});
  *out = p;
  return ZOSH_OK;
}
void zosh_program_free(zosh_program* p) { delete p; }
uint32_t zosh_program_num_ops(const zosh_program* p) { return p ? (uint32_t)p->ops.size() : 0; }
const zos_op* zosh_program_ops(const zosh_program* p) { return p && !p->ops.empty() ? p->ops.data() : nullptr; }
zos_status zosh_program_lower(const zosh_program* p, zos_ctx* ctx, uint32_t fuse_mode, uint32_t batch, zos_program** out) {
  if (!p) return ZOS_ERR_INVALID;
  return zos_program_create(ctx, p->ops.data(), (uint32_t)p->ops.size(), fuse_mode, batch, out);
}

}  // extern "C"
