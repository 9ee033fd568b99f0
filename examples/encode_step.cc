// The batched step of INTEGRATION.md section 3 from plain C++ (what a reference-side caller writes):
// one bi-predicted picture -- ME -> MC -> T/Q/recon -> deblocking -> padding -- through the C ABI of
// include/xvc_b200.h, on a synthetic picture cut into 32x32 CUs.
//
//   g++ -std=c++11 -I include examples/encode_step.cc -o encode_step -L xvc_b200 -lxvc_b200 -Wl,-rpath,$PWD/xvc_b200
//   ./encode_step [width height]        exit 0: step done, prints a checksum of the reconstruction
//                                       exit 3: no CUDA device (the library has no CPU fallback)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "xvc_b200.h"

namespace {
struct Picture {
  int w, h;
  std::vector<uint16_t> y, u, v;
  Picture(int w_, int h_) : w(w_), h(h_), y(size_t(w_) * h_), u(size_t(w_ / 2) * (h_ / 2)), v(size_t(w_ / 2) * (h_ / 2)) {}
  void Fill(int shift_x, int shift_y) {      // smooth texture, panned: something the search can find
    for (int r = 0; r < h; r++)
      for (int c = 0; c < w; c++)
        y[size_t(r) * w + c] = uint16_t(512 + 300 * std::sin(0.07 * (c + shift_x)) * std::cos(0.05 * (r + shift_y)) + ((c * 7 + r * 13) & 15));
    for (size_t i = 0; i < u.size(); i++) { u[i] = 480; v[i] = 540; }
  }
  const uint16_t *const *Planes(const uint16_t *p[3]) const { p[0] = y.data(); p[1] = u.data(); p[2] = v.data(); return p; }
};

int Fail(xvcb200_ctx *ctx, const char *what, int st) {
  std::fprintf(stderr, "%s failed (%d): %s\n", what, st, ctx ? xvcb200_ctx_error_string(ctx) : xvcb200_last_error_string());
  if (ctx) xvcb200_ctx_destroy(ctx);
  return st == XVCB200_NO_DEVICE ? 3 : 1;
}
}  // namespace

int main(int argc, char **argv) {
  const int width = argc > 2 ? std::atoi(argv[1]) : 256, height = argc > 2 ? std::atoi(argv[2]) : 128;
  const int bitdepth = 10, qp = 32;
  if (width < 32 || height < 32 || width % 32 || height % 32) { std::fprintf(stderr, "width and height must be multiples of 32\n"); return 2; }
  enum { ORIG, REF0, REF1, PRED, REC, LEV, NUM_SLOTS };
  xvcb200_ctx *ctx = nullptr;
  int st = xvcb200_ctx_create(&ctx, /*device*/ 0, width, height, bitdepth, /*4:2:0*/ 1, NUM_SLOTS);
  if (st != XVCB200_OK) return Fail(ctx, "xvcb200_ctx_create", st);

  Picture cur(width, height), r0(width, height), r1(width, height);
  cur.Fill(0, 0); r0.Fill(-3, 1); r1.Fill(3, -1);
  const ptrdiff_t strides[3] = {width, width / 2, width / 2};
  const uint16_t *p[3];
  xvcb200_upload_picture(ctx, ORIG, cur.Planes(p), strides);
  xvcb200_upload_picture(ctx, REF0, r0.Planes(p), strides);
  xvcb200_upload_picture(ctx, REF1, r1.Planes(p), strides);
  xvcb200_pad_border(ctx, REF0);           // references are padded once, when they are finished
  xvcb200_pad_border(ctx, REF1);

  std::vector<xvcb200_cu> cus;             // leaf CUs in coding order (CTUs in raster order, quadrants in z-order);
  for (int cy = 0; cy < height; cy += 64)  // mv[list] = the predictor of the search (zero here)
    for (int cx = 0; cx < width; cx += 64)
      for (int q = 0; q < 4; q++) {
        const int x = cx + 32 * (q & 1), y = cy + 32 * (q >> 1);
        if (x >= width || y >= height) continue;
        xvcb200_cu cu = {};
        cu.x = int16_t(x); cu.y = int16_t(y); cu.w = 32; cu.h = 32;
        cu.depth = 1; cu.qp = int8_t(qp); cu.ref_idx[0] = -1; cu.ref_idx[1] = -1;    // the step writes the chosen list back
        cus.push_back(cu);
      }
  const int n = int(cus.size());
  xvcb200_set_cus(ctx, cus.data(), n);

  xvcb200_picture_params prm = {};
  prm.orig_slot = ORIG; prm.pred_slot = PRED; prm.rec_slot = REC; prm.coeff_slot = LEV;
  for (int l = 0; l < 2; l++)
    for (int i = 0; i < 5; i++) prm.ref_slots[l][i] = -1;
  prm.ref_slots[0][0] = REF0; prm.ref_slots[1][0] = REF1;
  prm.ref_poc[0][0] = 0; prm.ref_poc[1][0] = 16;
  prm.num_ref[0] = prm.num_ref[1] = 1;
  prm.pic_type = 0;
  prm.search_range[0][0] = prm.search_range[1][0] = 64;
  prm.lambda_sqrt = std::sqrt(0.68 * std::pow(2.0, (qp - 12) / 3.0));      // Qp::GetLambdaSqrt of the picture
  prm.chroma_offset_table = 1; prm.deblock = 1; prm.pad = 1;
  std::vector<xvcb200_me_result> me(2 * size_t(n));
  std::vector<xvcb200_tu_result> tu(3 * size_t(n));
  st = xvcb200_encode_picture(ctx, &prm, me.data(), tu.data());            // asynchronous
  if (st != XVCB200_OK) return Fail(ctx, "xvcb200_encode_picture", st);

  Picture rec(width, height);
  uint16_t *out[3] = {rec.y.data(), rec.u.data(), rec.v.data()};
  xvcb200_download_picture(ctx, REC, out, strides);
  xvcb200_get_cus(ctx, cus.data(), n);                                     // chosen list, mv, cbf per CU
  st = xvcb200_sync(ctx);                                                  // sticky status of everything above
  if (st != XVCB200_OK) return Fail(ctx, "xvcb200_sync", st);

  uint64_t sum = 0;
  for (size_t i = 0; i < rec.y.size(); i++) sum = sum * 1099511628211ull + rec.y[i];
  int coded = 0;
  for (int i = 0; i < n; i++) coded += (cus[i].flags & XVCB200_CU_CBF_Y) != 0;
  std::printf("%dx%d: %d CUs, %d with luma coefficients, first CU mv L0 (%d,%d)/16 pel, recon checksum %016llx, %llu kernel launches\n",
              width, height, n, coded, cus[0].mv[0][0], cus[0].mv[0][1], (unsigned long long)sum,
              (unsigned long long)xvcb200_launch_count());
  xvcb200_ctx_destroy(ctx);
  return 0;
}
