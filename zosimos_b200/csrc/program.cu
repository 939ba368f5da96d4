// program.cu -- lowering of a High-like instruction stream to a schedule of fused kernels, and
// its stepwise execution.
//
// Replaces Program::lower_to / lower_to_impl (lib/zosimos/src/program.rs:1304-1651), the wgpu
// Encoder (lib/zosimos/src/program/encoder.rs) and the Low interpreter Host::step_inner
// (lib/zosimos/src/run.rs:1502-2365).  Where the reference emits, per operation, a decode pass, a
// render pass and an encode pass with a queue.submit each (program.rs:1475-1533), this planner
// chains per-pixel operations through their registers:
//
//   * a register written by a per-pixel op and read exactly once is never materialised; in
//     ZOS_FUSE_EXACT mode its declared texel quantisation (f16 texture + truncating pack, or the
//     native 8-bit rounding) is replayed in registers as a ZOS_STEP_REQUANT, so the bytes that
//     reach the outputs are the ones the unfused pass sequence produces;
//   * per-pixel producers of a composition's `above` operand become its source-side steps and
//     per-pixel consumers of a composition become its destination-side steps.
#include <string.h>

#include <vector>

#include "zos_internal.h"

using namespace zos;

namespace {

enum KKind { K_PIXEL, K_COMPOSE, K_COPY, K_GENERATE, K_BOX3, K_PALETTE, K_BUFFER_INIT, K_FROM_BUFFER, K_DYNAMIC };

struct Kernel {
  KKind kind;
  int src0 = -1, src1 = -1, dst = -1;  // registers (all materialised when the kernel runs)
  uint32_t nsteps = 0;
  zos_step steps[ZOS_MAX_STEPS];       // K_PIXEL
  zos_compose_params cp;               // K_COMPOSE (+ xc,yc of K_PALETTE in cp.inv)
  float gen[24];                       // K_GENERATE; K_BOX3 uses gen[0..8]
  uint32_t knob = 0;                   // knob id patched into this kernel's parameter block
  int knob_step = -1;                  // which step carries the knob-able matrix (K_PIXEL)
  std::vector<uint8_t> bytes;          // K_BUFFER_INIT: the initial content; K_DYNAMIC: the parameter block (both knob-able)
  zos_dynamic* dyn = nullptr;          // K_DYNAMIC: the compiled plugin (owned by the context's cache)
};

struct Reg {
  zos_desc desc;
  bool defined = false, is_input = false, is_output = false;
  bool bound = false, materialised = false;
  int uses = 0;
  zos_image img;
  zos_buf* owned = nullptr;
  bool has_pending = false;
  Kernel pending;
  bool is_buffer = false;  // a byte buffer register (ZOS_OP_BUFFER_INIT): `owned` is its storage, buf_bytes its size
  uint64_t buf_bytes = 0;
  uint64_t owned_bytes = 0;   // size of the storage the program itself provides for this register (0 = none)
  void* last_ptr = nullptr;   // where that storage was before zos_program_release_buffers (a recaptured graph is needed if it moves)
};

}  // namespace

struct zos_program {
  zos_ctx* ctx;
  uint32_t batch, fuse_mode;
  std::vector<Reg> regs;
  std::vector<Kernel> schedule;
  size_t pc = 0;
  bool running = false;
  // zos_program_run: the whole schedule as one CUDA graph, re-captured after a bind / knob change
  cudaGraphExec_t graph_exec = nullptr;
  bool graph_dirty = true, graph_broken = false;
  uint64_t runs = 0, graph_launches = 0;
  bool released = false;  // temporaries are parked in the context's arena (zos_program_release_buffers)
  // parameter blocks as planned, for every kernel a knob can patch: a cached program starts each launch from them
  // (a knob belongs to ONE Environment, run.rs:1292-1306; the next environment sees the initial content again)
  std::vector<std::pair<size_t, Kernel>> pristine;
  bool knobs_patched = false;
};

namespace {

zos_status alloc_reg(zos_program* p, int r) {
  Reg& R = p->regs[r];
  if (R.materialised) return ZOS_OK;
  R.materialised = true;
  if (R.is_input || R.bound) return ZOS_OK;  // storage comes from zos_program_bind
  zos_desc d = R.desc;
  if (d.block != ZOS_BLOCK_PIXEL) return fail(p->ctx, ZOS_ERR_UNSUPPORTED, "planar intermediate registers");
  d.row_stride = zos_aligned_row_stride(d.width, d.texel_stride);
  R.desc = d;
  uint64_t frame = d.row_stride * d.height;
  zos_status st = zos_buf_alloc(p->ctx, frame * p->batch, &R.owned);
  if (st != ZOS_OK) return st;
  R.owned_bytes = frame * p->batch;
  memset(&R.img, 0, sizeof R.img);
  R.img.desc = d;
  R.img.data = R.owned->ptr;
  R.img.batch_stride = p->batch > 1 ? frame : 0;
  return ZOS_OK;
}

zos_status flush(zos_program* p, int r) {  // make register r exist in memory
  Reg& R = p->regs[r];
  if (R.has_pending) {
    Kernel k = R.pending;
    R.has_pending = false;
    zos_status st = alloc_reg(p, r);
    if (st != ZOS_OK) return st;
    p->schedule.push_back(k);
    return ZOS_OK;
  }
  return alloc_reg(p, r);
}

bool append_requant(zos_program* p, zos_step* steps, uint32_t& n, const zos_desc& d) {
  if (p->fuse_mode != ZOS_FUSE_EXACT) return true;
  zos_texfmt f;
  if (zos_desc_texfmt(&d, &f) != ZOS_OK || f.storage > ZOS_STORAGE_FLOAT) return false;
  if (n >= ZOS_MAX_STEPS) return false;
  memset(&steps[n], 0, sizeof(zos_step));
  steps[n].kind = ZOS_STEP_REQUANT;
  steps[n].fmt = f;
  n++;
  return true;
}
bool append_steps(zos_step* steps, uint32_t& n, const zos_step* add, uint32_t nadd) {
  if (n + nadd > ZOS_MAX_STEPS) return false;
  for (uint32_t i = 0; i < nadd; i++) steps[n++] = add[i];
  return true;
}

bool same_chroma(const zos_desc& a, const zos_desc& b) {
  return a.block == b.block && a.bits == b.bits && a.parts == b.parts && a.color == b.color && a.transfer == b.transfer &&
         a.primaries == b.primaries && a.whitepoint == b.whitepoint;
}

zos_status check_reg(zos_program* p, int r, const char* what) {
  if (r < 0 || (size_t)r >= p->regs.size() || !p->regs[r].defined) return fail(p->ctx, ZOS_ERR_INVALID, "%s: bad register %d", what, r);
  return ZOS_OK;
}

zos_status plan(zos_program* p, const zos_op* ops, uint32_t nops) {
  zos_ctx* ctx = p->ctx;
  // registers are numbered by the op that defines them (`dst`), like Register(idx) in command.rs
  int maxreg = -1;
  for (uint32_t i = 0; i < nops; i++) maxreg = ops[i].dst > maxreg ? ops[i].dst : maxreg;
  p->regs.assign((size_t)(maxreg + 1), Reg());
  for (uint32_t i = 0; i < nops; i++)
    for (int s = 0; s < 2; s++)
      if (ops[i].src[s] >= 0 && ops[i].src[s] <= maxreg) p->regs[ops[i].src[s]].uses++;

  for (uint32_t i = 0; i < nops; i++) {
    const zos_op& op = ops[i];
    zos_status st;
    if (op.kind != ZOS_OP_OUTPUT) {
      if (op.dst < 0) return fail(ctx, ZOS_ERR_INVALID, "op %u: missing dst", i);
      Reg& D = p->regs[op.dst];
      if (D.defined) return fail(ctx, ZOS_ERR_INVALID, "op %u: register %d defined twice", i, op.dst);
      D.defined = true;
      D.desc = op.desc;
      zos_texfmt f;
      if (op.kind != ZOS_OP_BUFFER_INIT && (st = zos_desc_texfmt(&D.desc, &f)) != ZOS_OK)
        return fail(ctx, st, "op %u: register %d has no texture representation", i, op.dst);
    }
    switch (op.kind) {
      case ZOS_OP_INPUT:
        p->regs[op.dst].is_input = true;
        break;
      case ZOS_OP_OUTPUT:
        if ((st = check_reg(p, op.src[0], "output")) != ZOS_OK) return st;
        p->regs[op.src[0]].is_output = true;
        if ((st = flush(p, op.src[0])) != ZOS_OK) return st;
        break;
      case ZOS_OP_PIXEL: {
        if ((st = check_reg(p, op.src[0], "pixel op")) != ZOS_OK) return st;
        if ((st = validate_steps(ctx, op.steps, op.nsteps)) != ZOS_OK) return st;
        Reg& S = p->regs[op.src[0]];
        Reg& D = p->regs[op.dst];
        if (S.desc.width != D.desc.width || S.desc.height != D.desc.height) return fail(ctx, ZOS_ERR_TYPE, "op %u: size mismatch", i);
        bool fused = false;
        if (p->fuse_mode != ZOS_FUSE_NONE && S.has_pending && S.uses == 1 && !S.is_output && op.knob == 0 &&
            (S.pending.kind == K_PIXEL || S.pending.kind == K_COMPOSE)) {
          Kernel k = S.pending;
          zos_step* steps = k.kind == K_PIXEL ? k.steps : k.cp.dst_steps;
          uint32_t n = k.kind == K_PIXEL ? k.nsteps : k.cp.n_dst_steps;
          if (append_requant(p, steps, n, S.desc) && append_steps(steps, n, op.steps, op.nsteps)) {
            if (k.kind == K_PIXEL) k.nsteps = n; else k.cp.n_dst_steps = n;
            k.dst = op.dst;
            S.has_pending = false;
            D.pending = k;
            D.has_pending = true;
            fused = true;
          }
        }
        if (!fused) {
          if ((st = flush(p, op.src[0])) != ZOS_OK) return st;
          Kernel k;
          k.kind = K_PIXEL;
          k.src0 = op.src[0];
          k.dst = op.dst;
          k.nsteps = op.nsteps;
          for (uint32_t s = 0; s < op.nsteps; s++) k.steps[s] = op.steps[s];
          k.knob = op.knob;
          if (op.knob)
            for (uint32_t s = 0; s < op.nsteps; s++)
              if (op.steps[s].kind != ZOS_STEP_REQUANT && op.steps[s].kind != ZOS_STEP_F16) { k.knob_step = (int)s; break; }
          D.pending = k;
          D.has_pending = true;
        }
        if (p->fuse_mode == ZOS_FUSE_NONE || D.uses != 1) { if ((st = flush(p, op.dst)) != ZOS_OK) return st; }
        break;
      }
      case ZOS_OP_COMPOSE: {
        if ((st = check_reg(p, op.src[1], "compose (above)")) != ZOS_OK) return st;
        const bool has_below = op.src[0] >= 0;
        if (has_below && (st = check_reg(p, op.src[0], "compose (below)")) != ZOS_OK) return st;
        if ((st = validate_steps(ctx, op.compose.src_steps, op.compose.n_src_steps)) != ZOS_OK) return st;
        if ((st = validate_steps(ctx, op.compose.dst_steps, op.compose.n_dst_steps)) != ZOS_OK) return st;
        Reg& D = p->regs[op.dst];
        Kernel k;
        k.kind = K_COMPOSE;
        k.cp = op.compose;
        k.dst = op.dst;
        k.src0 = op.src[0];
        if (has_below) {
          Reg& B = p->regs[op.src[0]];
          if (B.desc.width != D.desc.width || B.desc.height != D.desc.height) return fail(ctx, ZOS_ERR_TYPE, "op %u: `below` and dst differ in size", i);
          if ((st = flush(p, op.src[0])) != ZOS_OK) return st;
        }
        Reg& A = p->regs[op.src[1]];
        k.src1 = op.src[1];
        if (p->fuse_mode != ZOS_FUSE_NONE && A.has_pending && A.uses == 1 && !A.is_output && A.pending.kind == K_PIXEL &&
            A.pending.knob == 0 && op.src[1] != op.src[0]) {
          // the producer chain of `above` becomes the source-side steps (applied per tap)
          zos_step steps[ZOS_MAX_STEPS];
          uint32_t n = 0;
          bool ok = append_steps(steps, n, A.pending.steps, A.pending.nsteps) && append_requant(p, steps, n, A.desc) &&
                    append_steps(steps, n, op.compose.src_steps, op.compose.n_src_steps);
          if (ok) {
            k.src1 = A.pending.src0;
            k.cp.n_src_steps = n;
            for (uint32_t s = 0; s < n; s++) k.cp.src_steps[s] = steps[s];
            A.has_pending = false;
          }
        }
        if (k.src1 == op.src[1] && (st = flush(p, op.src[1])) != ZOS_OK) return st;
        D.pending = k;
        D.has_pending = true;
        if (p->fuse_mode == ZOS_FUSE_NONE || D.uses != 1) { if ((st = flush(p, op.dst)) != ZOS_OK) return st; }
        break;
      }
      case ZOS_OP_COPY: {
        if ((st = check_reg(p, op.src[0], "copy")) != ZOS_OK) return st;
        Reg& S = p->regs[op.src[0]];
        Reg& D = p->regs[op.dst];
        if (S.desc.texel_stride != D.desc.texel_stride || S.desc.width != D.desc.width || S.desc.height != D.desc.height)
          return fail(ctx, ZOS_ERR_TYPE, "op %u: transmute needs equal texel size and image size (command.rs:1305-1313)", i);
        if ((st = flush(p, op.src[0])) != ZOS_OK) return st;
        if ((st = alloc_reg(p, op.dst)) != ZOS_OK) return st;
        Kernel k;
        k.kind = K_COPY; k.src0 = op.src[0]; k.dst = op.dst;
        p->schedule.push_back(k);
        break;
      }
      case ZOS_OP_BUFFER_INIT: {
        Reg& D = p->regs[op.dst];
        if (op.data_len == 0 || op.data_len > (1ull << 32)) return fail(ctx, ZOS_ERR_INVALID, "op %u: buffer of %llu bytes", i, (unsigned long long)op.data_len);
        D.is_buffer = true; D.buf_bytes = op.data_len; D.materialised = true;
        D.owned_bytes = (op.data_len + 255) & ~255ull;
        if ((st = zos_buf_alloc(ctx, D.owned_bytes, &D.owned)) != ZOS_OK) return st;
        Kernel k;
        k.kind = K_BUFFER_INIT; k.dst = op.dst; k.knob = op.knob;
        k.bytes.assign((size_t)op.data_len, 0);
        if (op.data) memcpy(k.bytes.data(), op.data, (size_t)op.data_len);
        p->schedule.push_back(k);
        break;
      }
      case ZOS_OP_DYNAMIC: {
        if (!op.source) return fail(ctx, ZOS_ERR_INVALID, "op %u: dynamic operator without source", i);
        if (p->batch != 1) return fail(ctx, ZOS_ERR_UNSUPPORTED, "op %u: dynamic operators in a batched program", i);
        for (int s2 = 0; s2 < 2; s2++)
          if (op.src[s2] >= 0) {
            if ((st = check_reg(p, op.src[s2], "dynamic operand")) != ZOS_OK) return st;
            if (p->regs[op.src[s2]].is_buffer) return fail(ctx, ZOS_ERR_TYPE, "op %u: dynamic operands are images (CommandError::INVALID_CALL)", i);
            if ((st = flush(p, op.src[s2])) != ZOS_OK) return st;
          }
        if ((st = alloc_reg(p, op.dst)) != ZOS_OK) return st;
        Kernel k;
        k.kind = K_DYNAMIC; k.src0 = op.src[0]; k.src1 = op.src[1]; k.dst = op.dst; k.knob = op.knob;
        if (op.data && op.data_len) k.bytes.assign((const uint8_t*)op.data, (const uint8_t*)op.data + op.data_len);
        if ((st = zos_dynamic_create(ctx, op.source, &k.dyn)) != ZOS_OK) return st;
        p->schedule.push_back(k);
        break;
      }
      case ZOS_OP_FROM_BUFFER: {
        if ((st = check_reg(p, op.src[0], "from_buffer")) != ZOS_OK) return st;
        Reg& S = p->regs[op.src[0]];
        Reg& D = p->regs[op.dst];
        if (!S.is_buffer) return fail(ctx, ZOS_ERR_TYPE, "op %u: from_buffer needs a buffer register (CommandError::TYPE_ERR)", i);
        if (D.desc.block != ZOS_BLOCK_PIXEL) return fail(ctx, ZOS_ERR_UNSUPPORTED, "op %u: from_buffer of a planar image", i);
        if ((st = alloc_reg(p, op.dst)) != ZOS_OK) return st;
        if (S.buf_bytes < D.desc.row_stride * D.desc.height) return fail(ctx, ZOS_ERR_INVALID, "op %u: buffer smaller than the aligned image (command.rs:955-961)", i);
        if (p->batch != 1) return fail(ctx, ZOS_ERR_UNSUPPORTED, "op %u: from_buffer in a batched program", i);
        Kernel k;
        k.kind = K_FROM_BUFFER; k.src0 = op.src[0]; k.dst = op.dst;
        p->schedule.push_back(k);
        break;
      }
      case ZOS_OP_GENERATE: {
        if (op.src[0] >= 0) {  // WithBuffer: the parameter block lives in a device buffer
          if ((st = check_reg(p, op.src[0], "generate (with_buffer)")) != ZOS_OK) return st;
          if (!p->regs[op.src[0]].is_buffer || p->regs[op.src[0]].buf_bytes < 96)
            return fail(ctx, ZOS_ERR_TYPE, "op %u: with_buffer needs a buffer register of at least 96 bytes", i);
        }
        if ((st = alloc_reg(p, op.dst)) != ZOS_OK) return st;
        Kernel k;
        k.kind = K_GENERATE; k.dst = op.dst; k.knob = op.knob; k.src0 = op.src[0];
        memset(&k.cp, 0, sizeof k.cp);
        k.cp.map = op.compose.map;  // 1 = solid colour
        memcpy(k.gen, op.gen, sizeof k.gen);
        p->schedule.push_back(k);
        break;
      }
      case ZOS_OP_BOX3: {
        if ((st = check_reg(p, op.src[0], "box3")) != ZOS_OK) return st;
        if ((st = flush(p, op.src[0])) != ZOS_OK) return st;
        if ((st = alloc_reg(p, op.dst)) != ZOS_OK) return st;
        Kernel k;
        k.kind = K_BOX3; k.src0 = op.src[0]; k.dst = op.dst; k.knob = op.knob;
        memcpy(k.gen, op.gen, 9 * sizeof(float));
        p->schedule.push_back(k);
        break;
      }
      case ZOS_OP_PALETTE: {
        if ((st = check_reg(p, op.src[0], "palette")) != ZOS_OK) return st;
        if ((st = check_reg(p, op.src[1], "palette indices")) != ZOS_OK) return st;
        if ((st = flush(p, op.src[0])) != ZOS_OK) return st;
        if ((st = flush(p, op.src[1])) != ZOS_OK) return st;
        if ((st = alloc_reg(p, op.dst)) != ZOS_OK) return st;
        Kernel k;
        k.kind = K_PALETTE; k.src0 = op.src[0]; k.src1 = op.src[1]; k.dst = op.dst;
        k.cp = op.compose;
        p->schedule.push_back(k);
        break;
      }
      default:
        return fail(ctx, ZOS_ERR_INVALID, "op %u: unknown kind %u", i, op.kind);
    }
  }
  // anything still pending was never consumed: dead code, dropped (the reference's liveness
  // analysis does the same, command.rs:2216-2291)
  return ZOS_OK;
}

zos_status run_kernel(zos_program* p, const Kernel& k) {
  zos_ctx* ctx = p->ctx;
  auto img = [&](int r) -> const zos_image* { return r >= 0 ? &p->regs[r].img : nullptr; };
  for (int r : {k.src0, k.src1, k.dst})
    if (r >= 0 && !p->regs[r].is_buffer && !p->regs[r].img.data) return fail(ctx, ZOS_ERR_STATE, "register %d is not bound (StartError::MissingKey)", r);
  switch (k.kind) {
    case K_PIXEL: return zos_pixel_chain(ctx, img(k.src0), img(k.dst), k.steps, k.nsteps, p->batch);
    case K_COMPOSE: return zos_compose(ctx, img(k.src0), img(k.src1), img(k.dst), &k.cp, p->batch);
    case K_GENERATE:
      if (k.src0 >= 0) return zos_generate_from_buffer(ctx, img(k.dst), (uint32_t)k.cp.map, p->regs[k.src0].owned, 0, p->batch);
      return zos_generate(ctx, img(k.dst), (uint32_t)k.cp.map, k.gen, p->batch);
    case K_DYNAMIC:
      return zos_dynamic_launch(ctx, k.dyn, img(k.dst), img(k.src0), img(k.src1), k.bytes.data(), k.bytes.size());
    case K_BUFFER_INIT:  // (the host bytes belong to the program: they outlive the asynchronous copy)
      return check_cuda(ctx, cudaMemcpyAsync(p->regs[k.dst].owned->ptr, k.bytes.data(), k.bytes.size(), cudaMemcpyHostToDevice, ctx->stream), "buffer_init");
    case K_FROM_BUFFER: {
      const zos_image* d = img(k.dst);
      return check_cuda(ctx, cudaMemcpyAsync(d->data, p->regs[k.src0].owned->ptr, d->desc.row_stride * d->desc.height, cudaMemcpyDeviceToDevice, ctx->stream), "from_buffer");
    }
    case K_BOX3: return zos_box3(ctx, img(k.src0), img(k.dst), k.gen, p->batch);
    case K_PALETTE: return zos_palette(ctx, img(k.src0), img(k.src1), img(k.dst), k.cp.inv, k.cp.inv + 4, p->batch);
    case K_COPY: {
      const zos_image* s = img(k.src0);
      const zos_image* d = img(k.dst);
      uint64_t row = (uint64_t)s->desc.width * s->desc.texel_stride;
      for (uint32_t f = 0; f < p->batch; f++) {
        cudaError_t e = cudaMemcpy2DAsync((uint8_t*)d->data + f * d->batch_stride, d->desc.row_stride,
                                          (const uint8_t*)s->data + f * s->batch_stride, s->desc.row_stride, row, s->desc.height,
                                          cudaMemcpyDeviceToDevice, ctx->stream);
        zos_status st = check_cuda(ctx, e, "High::Copy");
        if (st != ZOS_OK) return st;
      }
      return ZOS_OK;
    }
  }
  return ZOS_ERR_INVALID;
}

}  // namespace

extern "C" {

zos_status zos_program_create(zos_ctx* ctx, const zos_op* ops, uint32_t nops, uint32_t fuse_mode, uint32_t batch, zos_program** out) {
  if (!ctx || !out || (!ops && nops)) return ZOS_ERR_INVALID;
  *out = nullptr;
  if (fuse_mode > ZOS_FUSE_NONE) return fail(ctx, ZOS_ERR_INVALID, "bad fuse mode");
  if (batch == 0) return fail(ctx, ZOS_ERR_INVALID, "batch must be >= 1");
  zos_program* p = new zos_program();
  p->ctx = ctx;
  p->batch = batch;
  p->fuse_mode = fuse_mode;
  zos_status st = plan(p, ops, nops);
  if (st != ZOS_OK) { zos_program_destroy(p); return st; }
  for (size_t i = 0; i < p->schedule.size(); i++)
    if (p->schedule[i].knob) p->pristine.emplace_back(i, p->schedule[i]);
  *out = p;
  return ZOS_OK;
}

void zos_program_destroy(zos_program* p) {
  if (!p) return;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  for (Reg& r : p->regs)
    if (r.owned) zos_buf_free(p->ctx, r.owned);
  delete p;
}

zos_status zos_program_bind(zos_program* p, int32_t reg, const zos_image* image) {
  if (!p || !image) return ZOS_ERR_INVALID;
  zos_ctx* ctx = p->ctx;
  if (reg < 0 || (size_t)reg >= p->regs.size() || !p->regs[reg].defined) return fail(ctx, ZOS_ERR_INVALID, "bind: bad register %d", reg);
  Reg& R = p->regs[reg];
  if (!R.is_input && !R.is_output) return fail(ctx, ZOS_ERR_STATE, "bind: register %d is neither an input nor an output", reg);
  if (p->running) return fail(ctx, ZOS_ERR_STATE, "bind: program is running");
  const zos_desc& d = image->desc;
  if (d.width != R.desc.width || d.height != R.desc.height || !same_chroma(d, R.desc))
    return fail(ctx, ZOS_ERR_TYPE, "bind: descriptor mismatch for register %d (StartError::MismatchedDescriptor)", reg);
  if (!image->data) return fail(ctx, ZOS_ERR_INVALID, "bind: null image data");
  if (p->batch > 1 && image->batch_stride == 0) return fail(ctx, ZOS_ERR_INVALID, "bind: batch_stride required for batch > 1");
  if (R.owned) { zos_buf_free(ctx, R.owned); R.owned = nullptr; }
  R.owned_bytes = 0;  // the caller provides this register from now on
  if (R.bound && R.img.data == image->data && R.img.plane1 == image->plane1 && R.img.plane2 == image->plane2 &&
      R.img.batch_stride == image->batch_stride && R.img.desc.row_stride == image->desc.row_stride && R.img.chroma_stride == image->chroma_stride) {
    R.img = *image;  // the very same storage as last time: the captured graph stays valid (relaunch loops rebind every run)
    return ZOS_OK;
  }
  R.img = *image;
  R.bound = true;
  p->graph_dirty = true;
  return ZOS_OK;
}

// A cached program is launched again with other bindings: a register that is not bound this time goes back to what it was
// after planning -- an input waits for its image (StartError::MissingKey), anything else gets storage from the program.
zos_status zos_program_unbind(zos_program* p, int32_t reg) {
  if (!p) return ZOS_ERR_INVALID;
  zos_ctx* ctx = p->ctx;
  if (reg < 0 || (size_t)reg >= p->regs.size() || !p->regs[reg].defined) return fail(ctx, ZOS_ERR_INVALID, "unbind: bad register %d", reg);
  if (p->running) return fail(ctx, ZOS_ERR_STATE, "unbind: program is running");
  Reg& R = p->regs[reg];
  if (!R.bound) return ZOS_OK;
  R.bound = false;
  R.img.data = nullptr;
  p->graph_dirty = true;
  if (!R.is_input && R.materialised && !R.is_buffer) {
    zos_desc d = R.desc;
    d.row_stride = zos_aligned_row_stride(d.width, d.texel_stride);
    const uint64_t frame = d.row_stride * d.height;
    memset(&R.img, 0, sizeof R.img);
    R.img.desc = d;
    R.img.batch_stride = p->batch > 1 ? frame : 0;
    R.owned_bytes = frame * p->batch;
    R.last_ptr = nullptr;
    p->released = true;  // the next launch allocates it together with whatever else is parked
  }
  return ZOS_OK;
}

zos_status zos_program_reset_knobs(zos_program* p) {
  if (!p) return ZOS_ERR_INVALID;
  if (p->running) return fail(p->ctx, ZOS_ERR_STATE, "reset_knobs: program is running");
  if (!p->knobs_patched) return ZOS_OK;
  for (auto& pr : p->pristine) p->schedule[pr.first] = pr.second;
  p->knobs_patched = false;
  p->graph_dirty = true;
  return ZOS_OK;
}

zos_status zos_program_set_knob(zos_program* p, uint32_t knob, const void* data, uint64_t len) {
  if (!p || !data || knob == 0) return ZOS_ERR_INVALID;
  bool found = false;
  p->graph_dirty = true;
  p->knobs_patched = true;
  for (Kernel& k : p->schedule) {
    if (k.knob != knob) continue;
    found = true;
    const float* f = (const float*)data;
    if (k.kind == K_BUFFER_INIT || k.kind == K_DYNAMIC) {  // the whole content is the parameter block (tests/buffer.rs:67-118)
      if (len != k.bytes.size()) return fail(p->ctx, ZOS_ERR_INVALID, "knob %u: buffer holds %zu bytes", knob, k.bytes.size());
      memcpy(k.bytes.data(), data, (size_t)len);
      continue;
    }
    if (k.kind == K_GENERATE) {  // bilinear: 96 bytes (shaders/bilinear.rs:34-45); solid: 16 bytes (shaders/solid_rgb.rs:22-24)
      if (k.cp.map == ZOS_GEN_NORMAL2D) {  // vec2, mat2x2, float (shaders/distribution_normal2d.rs): 28 bytes, padded to 32
        if (len != 28 && len != 32) return fail(p->ctx, ZOS_ERR_INVALID, "knob %u: expected the 28/32-byte normal2d block", knob);
        memcpy(k.gen, f, 28);
      } else if (k.cp.map == ZOS_GEN_FRACTAL_NOISE) {  // vec2, float, float, uint (shaders/fractal_noise.rs:84-91): 20 bytes, padded to 24
        if (len != 20 && len != 24) return fail(p->ctx, ZOS_ERR_INVALID, "knob %u: expected the 20/24-byte fractal noise block", knob);
        memcpy(k.gen, f, 16);
        uint32_t oct; memcpy(&oct, (const uint8_t*)data + 16, 4);
        if (oct > 64) return fail(p->ctx, ZOS_ERR_INVALID, "knob %u: more than 64 octaves", knob);
        k.gen[4] = (float)oct;
      } else if (len == 96) memcpy(k.gen, f, 96);
      else if (len == 16) { memset(k.gen, 0, sizeof k.gen); memcpy(k.gen, f, 16); }
      else return fail(p->ctx, ZOS_ERR_INVALID, "knob %u: expected 96 or 16 bytes", knob);
    } else if ((k.kind == K_PIXEL && k.knob_step >= 0) || k.kind == K_BOX3) {
      // mat3 as 3 padded columns (color_matrix.rs:151-161): 48 bytes
      if (len != 48) return fail(p->ctx, ZOS_ERR_INVALID, "knob %u: expected a 48-byte std140 mat3", knob);
      float* m = k.kind == K_BOX3 ? k.gen : k.steps[k.knob_step].m;
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) m[3 * r + c] = f[4 * c + r];
    } else {
      return fail(p->ctx, ZOS_ERR_UNSUPPORTED, "knob %u: this operation has no patchable parameter block", knob);
    }
  }
  return found ? ZOS_OK : fail(p->ctx, ZOS_ERR_INVALID, "unknown knob %u", knob);
}

zos_status zos_program_launch(zos_program* p) {
  if (!p) return ZOS_ERR_INVALID;
  for (size_t r = 0; r < p->regs.size(); r++) {
    Reg& R = p->regs[r];
    if (R.defined && R.is_input && R.materialised && !R.bound) return fail(p->ctx, ZOS_ERR_STATE, "input register %zu is not bound (StartError::MissingKey)", r);
  }
  if (p->released) {
    zos_status st = zos_program_recover_buffers(p, nullptr, nullptr);
    if (st != ZOS_OK) return st;
  }
  p->pc = 0;
  p->running = !p->schedule.empty();  // nothing to step through (an output taken straight from an input): not in flight
  return ZOS_OK;
}

// Retire::retire_buffers / Environment::recover_buffers (run.rs:2876-2942, 1312-1347): between two launches the storage the
// program provides for its registers can wait in the context's arena, where any other program may take it.
zos_status zos_program_release_buffers(zos_program* p, uint64_t* bytes, uint32_t* count) {
  if (!p) return ZOS_ERR_INVALID;
  if (p->running) return fail(p->ctx, ZOS_ERR_STATE, "release_buffers: program is running");
  uint64_t b = 0;
  uint32_t n = 0;
  for (Reg& R : p->regs) {
    if (!R.owned) continue;
    R.last_ptr = R.owned->ptr;
    b += R.owned_bytes;
    n++;
    zos_buf_free(p->ctx, R.owned);
    R.owned = nullptr;
    if (!R.is_buffer) R.img.data = nullptr;
  }
  p->released = true;
  if (bytes) *bytes = b;
  if (count) *count = n;
  return ZOS_OK;
}
zos_status zos_program_recover_buffers(zos_program* p, uint64_t* bytes_reused, uint64_t* bytes_allocated) {
  if (!p) return ZOS_ERR_INVALID;
  uint64_t reused = 0, fresh = 0;
  for (Reg& R : p->regs) {
    if (R.owned || R.owned_bytes == 0) continue;
    const uint64_t before = p->ctx->arena.reuses;
    zos_status st = zos_buf_alloc(p->ctx, R.owned_bytes, &R.owned);
    if (st != ZOS_OK) return st;
    (p->ctx->arena.reuses != before ? reused : fresh) += R.owned_bytes;
    if (!R.is_buffer) R.img.data = R.owned->ptr;
    if (R.owned->ptr != R.last_ptr) p->graph_dirty = true;
  }
  p->released = false;
  if (bytes_reused) *bytes_reused = reused;
  if (bytes_allocated) *bytes_allocated = fresh;
  return ZOS_OK;
}
zos_status zos_program_resources(const zos_program* p, zos_program_stats* out) {
  if (!p || !out) return ZOS_ERR_INVALID;
  memset(out, 0, sizeof *out);
  out->kernels = (uint32_t)p->schedule.size();
  for (const Reg& R : p->regs)
    if (R.owned_bytes) { out->temp_buffers++; out->temp_bytes += R.owned_bytes; }
  out->released = p->released ? 1 : 0;
  out->runs = p->runs;
  out->graph_launches = p->graph_launches;
  return ZOS_OK;
}

zos_status zos_program_step(zos_program* p, uint32_t max_kernels, int32_t* still_running) {
  if (!p) return ZOS_ERR_INVALID;
  if (!p->running) return fail(p->ctx, ZOS_ERR_STATE, "step: program is not running (StepError::ProgramEnd)");
  for (uint32_t n = 0; n < max_kernels && p->pc < p->schedule.size(); n++) {
    zos_status st = run_kernel(p, p->schedule[p->pc]);
    if (st != ZOS_OK) { p->running = false; return st; }
    p->pc++;
  }
  if (p->pc >= p->schedule.size()) p->running = false;
  if (still_running) *still_running = p->running ? 1 : 0;
  return ZOS_OK;
}

// Executable reuse (run.rs:1016,1283-1347; tests/loop.rs): run the whole schedule again without
// re-planning.  With ZOS_RUN_GRAPH the kernels of the schedule are captured once into a CUDA graph and
// relaunched as ONE submission; binding another image or patching a knob re-captures on the next run.
zos_status zos_program_run(zos_program* p, uint32_t flags) {
  if (!p) return ZOS_ERR_INVALID;
  zos_ctx* ctx = p->ctx;
  if (p->running) return fail(ctx, ZOS_ERR_STATE, "run: a stepped execution of this program is in flight");
  zos_status st = zos_program_launch(p);
  if (st != ZOS_OK) return st;
  p->running = false;
  const bool want_graph = (flags & ZOS_RUN_GRAPH) && !p->graph_broken && p->schedule.size() > 1 && p->runs > 0;  // first run is eager: it also configures the kernels
  p->runs++;
  if (want_graph && (p->graph_dirty || !p->graph_exec)) {
    if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
    const uint64_t launches_before = ctx->launches;
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      for (const Kernel& k : p->schedule)
        if ((st = run_kernel(p, k)) != ZOS_OK) break;
      e = cudaStreamEndCapture(ctx->stream, &graph);
      if (e == cudaSuccess && st == ZOS_OK) e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
    }
    ctx->launches = launches_before;
    if (e != cudaSuccess || st != ZOS_OK || !p->graph_exec) {  // something in the schedule cannot be captured: stay eager from now on
      cudaGetLastError();
      if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
      p->graph_broken = true;
    } else {
      p->graph_dirty = false;
    }
  }
  if (want_graph && p->graph_exec && !p->graph_dirty) {
    ctx->launches += p->schedule.size();
    p->graph_launches++;
    return check_cuda(ctx, cudaGraphLaunch(p->graph_exec, ctx->stream), "cudaGraphLaunch");
  }
  for (const Kernel& k : p->schedule)
    if ((st = run_kernel(p, k)) != ZOS_OK) return st;
  return ZOS_OK;
}

uint64_t zos_program_graph_launches(const zos_program* p) { return p ? p->graph_launches : 0; }

uint32_t zos_program_kernel_count(const zos_program* p) { return p ? (uint32_t)p->schedule.size() : 0; }

zos_status zos_program_register_image(const zos_program* p, int32_t reg, zos_image* out) {
  if (!p || !out) return ZOS_ERR_INVALID;
  if (reg < 0 || (size_t)reg >= p->regs.size() || !p->regs[reg].defined) return ZOS_ERR_INVALID;
  const Reg& R = p->regs[reg];
  if (!R.img.data) return ZOS_ERR_STATE;
  *out = R.img;
  return ZOS_OK;
}

}  // extern "C"
