/* Plain C caller of include/zosimos_host.h (no GPU needed: the op builder, invoke and link are host code).
 * Mirrors tests/generic.rs of the reference: a generic callee (palette look-up through a ramp), declared as a
 * function of `main`, invoked with the type of main's input, linked; the linked stream must equal the one of
 * the same pipeline written without the function.  Exit status 0 = all checks passed. */
#include <stdio.h>
#include <string.h>

#include "zosimos_host.h"

#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "%s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, zosh_last_error()); return 1; } } while (0)

static zos_desc rgba8(uint32_t w, uint32_t h, uint32_t color, uint32_t transfer) {
  zos_desc d;
  memset(&d, 0, sizeof d);
  d.width = w; d.height = h;
  d.block = ZOS_BLOCK_PIXEL; d.bits = ZOS_BITS_UINT8X4; d.parts = ZOS_PARTS_RGBA;
  d.texel_stride = zos_bits_bytes(d.bits);
  d.row_stride = zos_aligned_row_stride(w, d.texel_stride);
  d.color = color; d.transfer = transfer; d.primaries = ZOS_PRIM_BT709; d.whitepoint = ZOS_WP_D65;
  return d;
}

static int pipeline(zosh_cb* cb, int32_t image, const zos_desc* idx_desc, int32_t* result) {
  float ramp[24] = {0};
  const float xc[4] = {0, 1, 0, 0}, yc[4] = {1, 0, 0, 0};
  int32_t idx;
  ramp[20] = 1; ramp[21] = 1; /* uv_max = (1, 1, 0, 0) */
  if (zosh_cb_bilinear(cb, idx_desc, ramp, &idx) != ZOSH_OK) return 1;
  return zosh_cb_palette(cb, image, idx, xc, yc, result) != ZOSH_OK;
}

int main(void) {
  const zos_desc srgb = rgba8(512, 512, ZOS_COLOR_RGB, ZOS_TRANSFER_SRGB);
  const zos_desc idx_desc = rgba8(256, 128, ZOS_COLOR_SCALARS, ZOS_TRANSFER_LINEAR);
  zosh_cb *callee = zosh_cb_new(), *main_cb = zosh_cb_new(), *direct = zosh_cb_new();
  zosh_signature* sig = NULL;
  zosh_program *linked = NULL, *plain = NULL;
  int32_t var, in, res, out, f, image, results[4];
  uint32_t nres = 0, i;
  zos_desc got;

  /* the generic callee */
  CHECK(zosh_cb_generic(callee, &var) == ZOSH_OK && var == 0);
  CHECK(zosh_cb_input_generic(callee, var, &in) == ZOSH_OK);
  CHECK(zosh_cb_describe(callee, in, &got) != ZOSH_OK); /* no type before it is bound */
  CHECK(pipeline(callee, in, &idx_desc, &res) == 0);
  CHECK(zosh_cb_output(callee, res, &out) == ZOSH_OK);
  CHECK(zosh_cb_computed_signature(callee, &sig) == ZOSH_OK);
  CHECK(zosh_signature_num_generics(sig) == 1 && zosh_signature_num_inputs(sig) == 1 && zosh_signature_num_outputs(sig) == 1);
  CHECK(zosh_compile(callee, &linked) == ZOSH_ERR_UNIMPLEMENTED);

  /* main: declares the function, invokes it with the type of its input */
  CHECK(zosh_cb_function(main_cb, sig, &f) == ZOSH_OK && zosh_cb_num_functions(main_cb) == 1);
  CHECK(zosh_cb_input(main_cb, &srgb, &image) == ZOSH_OK);
  CHECK(zosh_cb_describe(main_cb, image, &got) == ZOSH_OK);
  CHECK(zosh_cb_invoke(main_cb, f, NULL, 0, &image, 1, results, 4, &nres) == ZOSH_ERR_TYPE);      /* INVALID_CALL */
  CHECK(zosh_cb_invoke(main_cb, f, &idx_desc, 1, &image, 1, results, 4, &nres) == ZOSH_ERR_TYPE); /* wrong bound type */
  CHECK(zosh_cb_invoke(main_cb, f + 1, &got, 1, &image, 1, results, 4, &nres) == ZOSH_ERR_OTHER); /* BAD_REGISTER */
  CHECK(zosh_cb_num_ops(main_cb) == 1);
  CHECK(zosh_cb_invoke(main_cb, f, &got, 1, &image, 1, results, 4, &nres) == ZOSH_OK && nres == 1);
  CHECK(zosh_cb_describe(main_cb, results[0], &got) == ZOSH_OK && got.width == 256 && got.height == 128 && got.color == ZOS_COLOR_RGB);
  CHECK(zosh_cb_output(main_cb, results[0], &out) == ZOSH_OK);
  {
    const zosh_cb* fns[1];
    const uint32_t good[1] = {1}, bad[1] = {2}, per[2] = {1, 0};
    fns[0] = callee;
    CHECK(zosh_link(main_cb, NULL, 0, fns, 1, bad, per, &linked) == ZOSH_ERR_OTHER);
    fns[0] = direct;
    CHECK(zosh_link(main_cb, NULL, 0, fns, 1, good, per, &linked) == ZOSH_ERR_TYPE); /* not the invoked function */
    fns[0] = callee;
    CHECK(zosh_link(main_cb, NULL, 0, fns, 1, good, per, &linked) == ZOSH_OK);
  }

  /* the same pipeline without the function */
  CHECK(zosh_cb_input(direct, &srgb, &image) == ZOSH_OK);
  CHECK(pipeline(direct, image, &idx_desc, &res) == 0);
  CHECK(zosh_cb_output(direct, res, &out) == ZOSH_OK);
  CHECK(zosh_compile(direct, &plain) == ZOSH_OK);
  CHECK(zosh_program_num_ops(linked) == zosh_program_num_ops(plain) && zosh_program_num_ops(plain) == 4);
  for (i = 0; i < zosh_program_num_ops(plain); i++) {
    const zos_op *a = zosh_program_ops(linked) + i, *b = zosh_program_ops(plain) + i;
    CHECK(a->kind == b->kind && a->src[0] == b->src[0] && a->src[1] == b->src[1] && a->dst == b->dst);
    CHECK(memcmp(&a->desc, &b->desc, sizeof a->desc) == 0 && memcmp(a->gen, b->gen, sizeof a->gen) == 0);
    CHECK(memcmp(&a->compose, &b->compose, sizeof a->compose) == 0);
  }
  zosh_program_free(linked); zosh_program_free(plain);
  zosh_signature_free(sig);
  zosh_cb_free(callee); zosh_cb_free(main_cb); zosh_cb_free(direct);
  puts("host_generic ok");
  return 0;
}
