// TEST INFRASTRUCTURE ONLY.  C-ABI shim over the UNMODIFIED reference (divideon/xvc),
// compiled together with the reference's own sources into oracle/_ref/libxvcref.so by
// oracle/Makefile.  It exists to (1) pin the C restatement in xvc_oracle.c against the
// real reference, (2) generate the golden vectors under tests/golden/, and (3) time the
// reference's CPU implementation of the hot path (bench.py --impl reference).
// Nothing in xvc_b200/ links or loads this.
//
// The shim only drives reference classes; it contains no codec arithmetic of its own.
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <set>
#include <string>
#include <vector>
#include <atomic>
#include <thread>

// The hot-path leaf functions are private/protected members in the reference.
#define private public
#define protected public
#include "xvc_common_lib/coding_unit.h"
#include "xvc_common_lib/deblocking_filter.h"
#include "xvc_common_lib/inter_prediction.h"
#include "xvc_common_lib/intra_prediction.h"
#include "xvc_common_lib/picture_data.h"
#include "xvc_common_lib/quantize.h"
#include "xvc_common_lib/segment_header.h"
#include "xvc_common_lib/simd_cpu.h"
#include "xvc_common_lib/transform.h"
#include "xvc_common_lib/transform_data.h"
#include "xvc_common_lib/yuv_pic.h"
#include "xvc_enc_lib/cu_writer.h"
#include "xvc_enc_lib/encoder.h"
#include "xvc_enc_lib/encoder_settings.h"
#include "xvc_enc_lib/picture_encoder.h"
#include "xvc_enc_lib/syntax_writer.h"
#include "xvc_enc_lib/xvcenc.h"
#include "xvc_common_lib/reference_list_sorter.h"
#include "xvc_enc_lib/encoder_simd_functions.h"
#include "xvc_enc_lib/inter_search.h"
#include "xvc_enc_lib/inter_tz_search.h"
#include "xvc_enc_lib/rdo_quant.h"
#include "xvc_enc_lib/sample_metric.h"
#include "xvc_enc_lib/transform_encoder.h"
// InterSearch::SubpelSearch / GetSubpelDist / SearchRefIdx are member templates defined in
// the reference's inter_search.cc and fully inlined there, so they cannot be linked from
// outside.  The unmodified source file is compiled as part of THIS translation unit instead
// (and left out of the object list of libxvcref.so, see Makefile).
#include "xvc_enc_lib/inter_search.cc"   // NOLINT(build/include)
#undef private
#undef protected

#include <dlfcn.h>

#include <cstddef>
#include <mutex>

#include "../include/xvc_b200.h"   // shared plain-C descriptor structs + the table structs

using namespace xvc;  // NOLINT

namespace {

// Layout contract of INTEGRATION.md section 1: the two table structs of include/xvc_b200.h are the
// reference's InterPrediction::SimdFunc / SampleMetric::SimdFunc member for member, so a maintainer
// registers the CUDA entries with one call on the reference's own table object.
#define XVCB_SAME_MEMBER(ours, theirs, m)                                                                   \
  static_assert(offsetof(ours, m) == offsetof(theirs, m) && sizeof(((ours *)0)->m) == sizeof(((theirs *)0)->m), \
                "table member " #m)
static_assert(sizeof(xvcb200_inter_prediction_simd_func) == sizeof(InterPrediction::SimdFunc), "InterPrediction::SimdFunc layout");
static_assert(sizeof(xvcb200_sample_metric_simd_func) == sizeof(SampleMetric::SimdFunc), "SampleMetric::SimdFunc layout");
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, add_avg);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_copy_bipred);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_h_sample_sample);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_h_sample_short);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_v_sample_sample);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_v_sample_short);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_v_short_sample);
XVCB_SAME_MEMBER(xvcb200_inter_prediction_simd_func, InterPrediction::SimdFunc, filter_v_short_short);
XVCB_SAME_MEMBER(xvcb200_sample_metric_simd_func, SampleMetric::SimdFunc, sad_sample_sample);
XVCB_SAME_MEMBER(xvcb200_sample_metric_simd_func, SampleMetric::SimdFunc, sad_short_sample);
XVCB_SAME_MEMBER(xvcb200_sample_metric_simd_func, SampleMetric::SimdFunc, ssd_sample_sample);
XVCB_SAME_MEMBER(xvcb200_sample_metric_simd_func, SampleMetric::SimdFunc, ssd_short_sample);
XVCB_SAME_MEMBER(xvcb200_sample_metric_simd_func, SampleMetric::SimdFunc, ssd_short_short);

// use_simd: 0 = the reference's C entries, 1 = its runtime SIMD entries, 2 = the C entries
// overwritten by libxvc_b200's Register functions (the drop-in of INTEGRATION.md section 1: the
// reference's own classes then run with the CUDA entries behind their tables).  The library is
// found through $XVCB200_LIB and resolved at run time, so libxvcref.so itself never links CUDA.
const EncoderSimdFunctions &Simd(int use_simd, int bitdepth) {
  // index: [use_simd][bitdepth-8]
  static std::unique_ptr<EncoderSimdFunctions> tables[3][9];
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  auto &slot = tables[use_simd < 0 || use_simd > 2 ? 0 : use_simd][bitdepth - 8];
  if (!slot) {
    std::set<CpuCapability> caps;
    if (use_simd == 1) caps = SimdCpu::GetRuntimeCapabilities();
    slot.reset(new EncoderSimdFunctions(caps, bitdepth));
    if (use_simd == 2) {
      const char *path = getenv("XVCB200_LIB");
      void *lib = dlopen(path ? path : "libxvc_b200.so", RTLD_NOW | RTLD_GLOBAL);
      typedef void (*RegInter)(xvcb200_inter_prediction_simd_func *);
      typedef void (*RegMetric)(int, xvcb200_sample_metric_simd_func *);
      RegInter reg_inter = lib ? reinterpret_cast<RegInter>(dlsym(lib, "xvcb200_register_inter_prediction")) : nullptr;
      RegMetric reg_metric = lib ? reinterpret_cast<RegMetric>(dlsym(lib, "xvcb200_register_sample_metric")) : nullptr;
      if (!reg_inter || !reg_metric) {
        fprintf(stderr, "ref_shim: cannot load the CUDA tables from %s: %s\n", path ? path : "libxvc_b200.so", dlerror());
        abort();
      }
      reg_inter(reinterpret_cast<xvcb200_inter_prediction_simd_func *>(&slot->inter_prediction));
      reg_metric(bitdepth, reinterpret_cast<xvcb200_sample_metric_simd_func *>(&slot->sample_metric));
    }
  }
  return *slot;
}

int Log2(int v) { return util::SizeToLog2(v); }

// Dynamic-chunk parallel loop over [0,n) with per-thread state built by make_state().
// Restrictions is thread_local in the reference and default-constructs to the
// unrestricted mode in every new thread, which is what this path uses.
template <typename State, typename MakeState, typename Body>
void ParallelFor(int n, int threads, MakeState make_state, Body body) {
  if (threads < 1) threads = 1;
  std::atomic<int> next(0);
  auto worker = [&]() {
    std::unique_ptr<State> st(make_state());
    for (;;) {
      int begin = next.fetch_add(8);
      if (begin >= n) break;
      int end = std::min(n, begin + 8);
      for (int i = begin; i < end; i++) body(st.get(), i);
    }
  };
  // always on fresh threads, also for threads == 1: the calling thread's Restrictions may have been
  // changed by an encoder / decoder that ran on it before (the conformance tests do)
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(worker);
  for (auto &t : pool) t.join();
}

}  // namespace

extern "C" {

// How many entries of the use_simd tables differ from the reference's C entries (0 for use_simd = 0).
int xref_table_entries_replaced(int use_simd, int bitdepth) {
  const EncoderSimdFunctions &c = Simd(0, bitdepth), &t = Simd(use_simd, bitdepth);
  int n = 0;
  const void *const *a = reinterpret_cast<const void *const *>(&c.inter_prediction);
  const void *const *b = reinterpret_cast<const void *const *>(&t.inter_prediction);
  for (size_t i = 0; i < sizeof(c.inter_prediction) / sizeof(void *); i++) n += a[i] != b[i];
  a = reinterpret_cast<const void *const *>(&c.sample_metric);
  b = reinterpret_cast<const void *const *>(&t.sample_metric);
  for (size_t i = 0; i < sizeof(c.sample_metric) / sizeof(void *); i++) n += a[i] != b[i];
  return n;
}

// ---------------------------------------------------------------- leaf: metrics
int xref_sad(int kind, int use_simd, int bitdepth, int w, int h, const void *a, ptrdiff_t sa,
             const uint16_t *b, ptrdiff_t sb) {
  const auto &t = Simd(use_simd, bitdepth).sample_metric;
  if (kind == 0) return t.sad_sample_sample[Log2(w)](w, h, static_cast<const Sample *>(a), sa, b, sb);
  return t.sad_short_sample[Log2(w)](w, h, static_cast<const int16_t *>(a), sa, b, sb);
}

uint64_t xref_ssd(int kind, int use_simd, int bitdepth, int w, int h, const void *a, ptrdiff_t sa,
                  const void *b, ptrdiff_t sb) {
  const auto &t = Simd(use_simd, bitdepth).sample_metric;
  if (kind == 0)
    return t.ssd_sample_sample[Log2(w)](w, h, static_cast<const Sample *>(a), sa,
                                        static_cast<const Sample *>(b), sb);
  if (kind == 1)
    return t.ssd_short_sample[Log2(w)](w, h, static_cast<const int16_t *>(a), sa,
                                       static_cast<const Sample *>(b), sb);
  return t.ssd_short_short[Log2(w)](w, h, static_cast<const int16_t *>(a), sa,
                                    static_cast<const int16_t *>(b), sb);
}

// SampleMetric::Compare with a luma Qp (weight 1.0).  metric = MetricType numbering.
// first_short: src1 is int16_t (Residual) instead of Sample.
uint64_t xref_compare(int metric, int use_simd, int bitdepth, int comp, int qp, int w, int h,
                      int first_short, const void *a, ptrdiff_t sa, const uint16_t *b, ptrdiff_t sb) {
  SampleMetric m(Simd(use_simd, bitdepth).sample_metric, bitdepth, static_cast<MetricType>(metric));
  Qp q(qp, ChromaFormat::k420, bitdepth, 1.0, 1, 0, 0);
  if (first_short)
    return m.Compare(q, static_cast<YuvComponent>(comp), w, h, static_cast<const Residual *>(a), sa, b, sb);
  return m.Compare(q, static_cast<YuvComponent>(comp), w, h, static_cast<const Sample *>(a), sa, b, sb);
}

// ---------------------------------------------------------------- leaf: filters
// kind: 0 h_sample_sample 1 h_sample_short 2 v_sample_sample 3 v_sample_short
//       4 v_short_sample 5 v_short_short
void xref_filter(int kind, int chroma, int use_simd, int w, int h, int bitdepth, const int16_t *taps,
                 const void *src, ptrdiff_t ss, void *dst, ptrdiff_t ds) {
  const auto &t = Simd(use_simd, bitdepth).inter_prediction;
  const Sample *s_u = static_cast<const Sample *>(src);
  const int16_t *s_s = static_cast<const int16_t *>(src);
  Sample *d_u = static_cast<Sample *>(dst);
  int16_t *d_s = static_cast<int16_t *>(dst);
  switch (kind) {
    case 0: t.filter_h_sample_sample[chroma](w, h, bitdepth, taps, s_u, ss, d_u, ds); break;
    case 1: t.filter_h_sample_short[chroma](w, h, bitdepth, taps, s_u, ss, d_s, ds); break;
    case 2: t.filter_v_sample_sample[chroma](w, h, bitdepth, taps, s_u, ss, d_u, ds); break;
    case 3: t.filter_v_sample_short[chroma](w, h, bitdepth, taps, s_u, ss, d_s, ds); break;
    case 4: t.filter_v_short_sample[chroma](w, h, bitdepth, taps, s_s, ss, d_u, ds); break;
    case 5: t.filter_v_short_short[chroma](w, h, bitdepth, taps, s_s, ss, d_s, ds); break;
    default: assert(0);
  }
}

void xref_add_avg(int use_simd, int w, int h, int offset, int shift, int bitdepth, const int16_t *a,
                  intptr_t sa, const int16_t *b, intptr_t sb, uint16_t *dst, intptr_t ds) {
  Simd(use_simd, bitdepth).inter_prediction.add_avg[w > 2](w, h, offset, shift, bitdepth, a, sa, b, sb, dst, ds);
}

void xref_filter_copy_bipred(int use_simd, int bitdepth, int w, int h, int offset, int shift,
                             const uint16_t *ref, ptrdiff_t rs, int16_t *pred, ptrdiff_t ps) {
  Simd(use_simd, bitdepth).inter_prediction.filter_copy_bipred[w > 2](
      w, h, static_cast<int16_t>(offset), shift, ref, rs, pred, ps);
}

// InterPrediction::FilterLuma/FilterChroma (+Bipred) on a raw block.
void xref_interp(int chroma, int bipred, int use_simd, int w, int h, int bitdepth, int frac_x, int frac_y,
                 const uint16_t *ref, ptrdiff_t rs, void *pred, ptrdiff_t ps) {
  YuvPicture dummy(ChromaFormat::k420, 0, 0, bitdepth, false, 0, 0);
  InterPrediction ip(Simd(use_simd, bitdepth).inter_prediction, dummy, bitdepth);
  if (frac_x == 0 && frac_y == 0) {
    if (bipred) {
      DataBuffer<int16_t> out(static_cast<int16_t *>(pred), ps);
      ip.FilterCopyBipred(w, h, SampleBufferConst(ref, rs), &out);
    } else {
      SampleBuffer out(static_cast<Sample *>(pred), ps);
      out.CopyFrom(w, h, SampleBufferConst(ref, rs));
    }
    return;
  }
  if (!bipred) {
    if (!chroma) ip.FilterLuma(w, h, frac_x, frac_y, ref, rs, static_cast<Sample *>(pred), ps);
    else ip.FilterChroma(w, h, frac_x, frac_y, ref, rs, static_cast<Sample *>(pred), ps);
  } else {
    if (!chroma) ip.FilterLumaBipred(w, h, frac_x, frac_y, ref, rs, static_cast<int16_t *>(pred), ps);
    else ip.FilterChromaBipred(w, h, frac_x, frac_y, ref, rs, static_cast<int16_t *>(pred), ps);
  }
}

// ---------------------------------------------------------------- leaf: transform / quant
namespace {
struct TxCu {
  TxCu(int w, int h, int bitdepth, int comp, bool intra, int tx_hor, int tx_ver)
      : pic_data(ChromaFormat::k420, comp == 0 ? w : 2 * w, comp == 0 ? h : 2 * h, bitdepth) {
    pic_data.SetNalType(NalUnitType::kBipredictedPicture);
    cu = pic_data.CreateCu(CuTree::Primary, 0, 0, 0, comp == 0 ? w : 2 * w, comp == 0 ? h : 2 * h);
    cu->SetPredMode(intra ? PredictionMode::kIntra : PredictionMode::kInter);
    // SetTransformType(comp, t1, t2): t1 = type[0] (vertical), t2 = type[1] (horizontal)
    cu->SetTransformType(static_cast<YuvComponent>(comp), static_cast<TransformType>(tx_ver),
                         static_cast<TransformType>(tx_hor));
    cu->SetDcCoeffOnly(static_cast<YuvComponent>(comp), false);
  }
  PictureData pic_data;
  CodingUnit *cu;
};
}  // namespace

void xref_fwd_transform(int w, int h, int bitdepth, int comp, int intra, int tx_hor, int tx_ver,
                        const int16_t *resi, ptrdiff_t rs, int16_t *coeff, ptrdiff_t cs) {
  TxCu t(w, h, bitdepth, comp, intra != 0, tx_hor, tx_ver);
  ForwardTransform fwd(bitdepth);
  ResidualBuffer in(const_cast<int16_t *>(resi), rs);
  CoeffBuffer out(coeff, cs);
  fwd.Transform(*t.cu, static_cast<YuvComponent>(comp), in, &out);
}

void xref_inv_transform(int w, int h, int bitdepth, int comp, int intra, int tx_hor, int tx_ver, int dc_only,
                        const int16_t *coeff, ptrdiff_t cs, int16_t *resi, ptrdiff_t rs) {
  TxCu t(w, h, bitdepth, comp, intra != 0, tx_hor, tx_ver);
  t.cu->SetDcCoeffOnly(static_cast<YuvComponent>(comp), dc_only != 0);
  InverseTransform inv(bitdepth);
  CoeffBuffer in(const_cast<int16_t *>(coeff), cs);
  ResidualBuffer out(resi, rs);
  inv.Transform(*t.cu, static_cast<YuvComponent>(comp), in, &out);
}

void xref_transform_skip(int forward, int w, int h, int bitdepth, const int16_t *in, ptrdiff_t is,
                         int16_t *out, ptrdiff_t os) {
  if (forward) {
    ForwardTransform fwd(bitdepth);
    ResidualBuffer i(const_cast<int16_t *>(in), is);
    CoeffBuffer o(out, os);
    fwd.TransformSkip(w, h, i, &o);
  } else {
    InverseTransform inv(bitdepth);
    CoeffBuffer i(const_cast<int16_t *>(in), is);
    ResidualBuffer o(out, os);
    inv.TransformSkip(w, h, i, &o);
  }
}

// qp is the raw LUMA qp; comp selects the component (chroma qp via table 1, offsets 0).
int xref_quant_fast(int w, int h, int bitdepth, int comp, int qp, int intra_pic, int intra_cu, int intra_mode,
                    const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  TxCu t(w, h, bitdepth, comp, intra_cu != 0, 0, 0);
  if (intra_cu) t.cu->SetIntraModeLuma(static_cast<IntraMode>(intra_mode));
  EncoderSettings settings;
  RdoQuant q(bitdepth, settings);
  Qp qpo(qp, ChromaFormat::k420, bitdepth, 1.0, 1, 0, 0);
  return q.QuantFast(*t.cu, static_cast<YuvComponent>(comp), qpo,
                     intra_pic ? PicturePredictionType::kIntra : PicturePredictionType::kBi, in, is, out, os);
}

// RDOQ with FROZEN contexts (the definition the GPU kernel of a later round has to meet):
// RdoQuant::QuantRdo (rdo_quant.cc:203-446) with the context state every picture starts from -- a SyntaxWriter
// freshly initialised for the picture's qp and prediction type (Contexts::ResetStates through its constructor,
// syntax_writer.cc:35-41), never advanced: QuantRdo only reads the writer, so the result of a transform unit does
// not depend on the units coded before it and all units of a picture can be quantised at once.  lambda is the
// picture's lambda (Qp carries it: rdo_quant.cc uses qp.GetLambda()); qp is the raw LUMA qp, comp selects the
// component (chroma qp via table 1, offsets 0), like xref_quant_fast.
int xref_quant_rdo_frozen(int w, int h, int bitdepth, int comp, int qp, double lambda, int intra_pic, int intra_cu,
                          int intra_mode, const int16_t *in, ptrdiff_t is, int16_t *out, ptrdiff_t os) {
  TxCu t(w, h, bitdepth, comp, intra_cu != 0, 0, 0);
  if (intra_cu) t.cu->SetIntraModeLuma(static_cast<IntraMode>(intra_mode));
  EncoderSettings settings;
  RdoQuant q(bitdepth, settings);
  Qp qpo(qp, ChromaFormat::k420, bitdepth, lambda, 1, 0, 0);
  const PicturePredictionType pic_type = intra_pic ? PicturePredictionType::kIntra : PicturePredictionType::kBi;
  BitWriter bw;
  SyntaxWriter writer(qpo, pic_type, &bw);
  return q.QuantRdo(*t.cu, static_cast<YuvComponent>(comp), qpo, pic_type, writer, in, is, out, os);
}

void xref_dequant(int w, int h, int bitdepth, int comp, int qp, const int16_t *in, ptrdiff_t is,
                  int16_t *out, ptrdiff_t os) {
  Quantize q;
  Qp qpo(qp, ChromaFormat::k420, bitdepth, 1.0, 1, 0, 0);
  q.Inverse(static_cast<YuvComponent>(comp), qpo, w, h, bitdepth, in, is, out, os);
}

void xref_qp_info(int qp, int bitdepth, double lambda, int table, int off_u, int off_v, xvcb200_qp *out) {
  Qp q(qp, ChromaFormat::k420, bitdepth, lambda, table, off_u, off_v);
  for (int c = 0; c < 3; c++) {
    out->qp_raw[c] = q.qp_raw_[c];
    out->qp_bitdepth[c] = q.qp_bitdepth_[c];
    out->distortion_weight[c] = q.distortion_weight_[c];
    out->lambda[c] = q.lambda_[c];
  }
  out->lambda_sqrt = q.lambda_sqrt_;
}

// Transform matrices for table extraction (tools/gen_tables.py).  kind: 0 DCT2 (6-bit),
// 1 DCT2 high, 2 DCT5, 3 DCT8, 4 DST1, 5 DST7.  Returns N*N row-major int16 or NULL.
const int16_t *xref_transform_matrix(int kind, int n) {
  typedef TransformData T;
  switch (kind) {
    case 0:
      switch (n) { case 4: return &T::kDct2Transform4[0][0]; case 8: return &T::kDct2Transform8[0][0];
                   case 16: return &T::kDct2Transform16[0][0]; case 32: return &T::kDct2Transform32[0][0]; }
      break;
    case 1:
      switch (n) { case 2: return &T::kDct2Transform2High[0][0]; case 4: return &T::kDct2Transform4High[0][0];
                   case 8: return &T::kDct2Transform8High[0][0]; case 16: return &T::kDct2Transform16High[0][0];
                   case 32: return &T::kDct2Transform32High[0][0]; case 64: return &T::kDct2Transform64High[0][0]; }
      break;
    case 2:
      switch (n) { case 4: return T::kDct5Transform4High; case 8: return T::kDct5Transform8High;
                   case 16: return T::kDct5Transform16High; case 32: return T::kDct5Transform32High;
                   case 64: return T::kDct5Transform64High; }
      break;
    case 3:
      switch (n) { case 4: return T::kDct8Transform4High; case 8: return T::kDct8Transform8High;
                   case 16: return T::kDct8Transform16High; case 32: return T::kDct8Transform32High;
                   case 64: return T::kDct8Transform64High; }
      break;
    case 4:
      switch (n) { case 4: return T::kDst1Transform4High; case 8: return T::kDst1Transform8High;
                   case 16: return T::kDst1Transform16High; case 32: return T::kDst1Transform32High;
                   case 64: return T::kDst1Transform64High; }
      break;
    case 5:
      switch (n) { case 4: return T::kDst7Transform4High; case 8: return T::kDst7Transform8High;
                   case 16: return T::kDst7Transform16High; case 32: return T::kDst7Transform32High;
                   case 64: return T::kDst7Transform64High; }
      break;
  }
  return nullptr;
}

// ---------------------------------------------------------------- picture session
struct xref_session {
  int width, height, bitdepth, pic_type, use_simd;
  int chroma_table, off_u, off_v;
  double lambda;
  int pic_qp;
  std::shared_ptr<PictureData> pic_data;
  std::shared_ptr<YuvPicture> orig;      // unpadded (picture_encoder.cc:46-48)
  std::shared_ptr<YuvPicture> rec;       // padded
  std::shared_ptr<YuvPicture> pred;      // padded layout, holds the prediction signal
  std::vector<std::shared_ptr<YuvPicture>> refs[2];
  std::vector<std::shared_ptr<PictureData>> ref_data[2];
  std::vector<int16_t> coeff[3];         // picture-shaped level planes (tight)
  EncoderSettings settings;
  SegmentHeader segment;
  std::vector<CodingUnit *> cus;
  bool inited = false;
};

static void CopyIn(YuvPicture *pic, const uint16_t *const planes[3]) {
  for (int c = 0; c < 3; c++) {
    YuvComponent comp = static_cast<YuvComponent>(c);
    int w = pic->GetWidth(comp), h = pic->GetHeight(comp);
    for (int y = 0; y < h; y++)
      std::memcpy(pic->GetSamplePtr(comp, 0, y), planes[c] + static_cast<size_t>(y) * w, sizeof(Sample) * w);
  }
}
static void CopyOut(const YuvPicture *pic, uint16_t *const planes[3]) {
  for (int c = 0; c < 3; c++) {
    YuvComponent comp = static_cast<YuvComponent>(c);
    int w = pic->GetWidth(comp), h = pic->GetHeight(comp);
    for (int y = 0; y < h; y++)
      std::memcpy(planes[c] + static_cast<size_t>(y) * w, pic->GetSamplePtr(comp, 0, y), sizeof(Sample) * w);
  }
}

xref_session *xref_session_create(int width, int height, int bitdepth, int pic_type, int qp, double lambda,
                                  int use_simd, int64_t poc, int sub_gop_length, int chroma_table,
                                  int off_u, int off_v) {
  xref_session *s = new xref_session();
  s->width = width; s->height = height; s->bitdepth = bitdepth; s->pic_type = pic_type;
  s->use_simd = use_simd; s->lambda = lambda; s->pic_qp = qp;
  s->chroma_table = chroma_table; s->off_u = off_u; s->off_v = off_v;
  s->pic_data.reset(new PictureData(ChromaFormat::k420, width, height, bitdepth));
  s->pic_data->SetNalType(pic_type == 0 ? NalUnitType::kBipredictedPicture :
                          pic_type == 1 ? NalUnitType::kPredictedPicture : NalUnitType::kIntraPicture);
  s->pic_data->SetPoc(static_cast<PicNum>(poc));
  s->pic_data->SetSubGopLength(static_cast<PicNum>(sub_gop_length));
  s->pic_data->SetTid(0);
  s->orig.reset(new YuvPicture(ChromaFormat::k420, width, height, bitdepth, false, 0, 0));
  s->rec.reset(new YuvPicture(ChromaFormat::k420, width, height, bitdepth, true, 0, 0));
  s->pred.reset(new YuvPicture(ChromaFormat::k420, width, height, bitdepth, true, 0, 0));
  for (int c = 0; c < 3; c++)
    s->coeff[c].assign(static_cast<size_t>(s->orig->GetWidth(YuvComponent(c))) * s->orig->GetHeight(YuvComponent(c)), 0);
  s->settings.Initialize(SpeedMode::kSlow);
  s->segment.chroma_qp_offset_table = chroma_table;
  s->segment.chroma_qp_offset_u = off_u;
  s->segment.chroma_qp_offset_v = off_v;
  s->segment.max_binary_split_depth = 3;
  return s;
}

void xref_session_destroy(xref_session *s) { delete s; }

void xref_session_set_orig(xref_session *s, const uint16_t *const planes[3]) { CopyIn(s->orig.get(), planes); }
void xref_session_set_rec(xref_session *s, const uint16_t *const planes[3]) { CopyIn(s->rec.get(), planes); }
void xref_session_get_rec(xref_session *s, uint16_t *const planes[3]) { CopyOut(s->rec.get(), planes); }
void xref_session_set_pred(xref_session *s, const uint16_t *const planes[3]) { CopyIn(s->pred.get(), planes); }
void xref_session_get_pred(xref_session *s, uint16_t *const planes[3]) { CopyOut(s->pred.get(), planes); }
void xref_session_get_coeff(xref_session *s, int16_t *const planes[3]) {
  for (int c = 0; c < 3; c++) std::memcpy(planes[c], s->coeff[c].data(), s->coeff[c].size() * sizeof(int16_t));
}

// Adds a reference picture: samples copied in, YuvPicture::PadBorder applied (the
// reference's own padding), registered in the ReferencePictureLists.
void xref_session_add_ref(xref_session *s, int list, int idx, int64_t poc, const uint16_t *const planes[3]) {
  std::shared_ptr<YuvPicture> pic(new YuvPicture(ChromaFormat::k420, s->width, s->height, s->bitdepth, true, 0, 0));
  CopyIn(pic.get(), planes);
  pic->PadBorder();
  std::shared_ptr<PictureData> pd(new PictureData(ChromaFormat::k420, s->width, s->height, s->bitdepth));
  pd->SetNalType(NalUnitType::kIntraPicture);
  pd->SetPoc(static_cast<PicNum>(poc));
  pd->SetTid(0);
  if (static_cast<int>(s->refs[list].size()) <= idx) { s->refs[list].resize(idx + 1); s->ref_data[list].resize(idx + 1); }
  s->refs[list][idx] = pic;
  s->ref_data[list][idx] = pd;
  s->pic_data->GetRefPicLists()->SetRefPic(static_cast<RefPicList>(list), idx, static_cast<PicNum>(poc), pd, pic, pic);
}

// Returns the padded reference (full allocation incl. border) for PadBorder parity.
void xref_session_get_ref_padded(xref_session *s, int list, int idx, int comp, uint16_t *out) {
  const YuvPicture *pic = s->refs[list][idx].get();
  YuvComponent c = static_cast<YuvComponent>(comp);
  int stride = static_cast<int>(pic->GetStride(c));
  int off_x = (stride - pic->GetWidth(c)) / 2;
  int off_y = (pic->GetTotalHeight(c) - pic->GetHeight(c)) / 2;
  const Sample *base = pic->GetSamplePtr(c, -off_x, -off_y);
  std::memcpy(out, base, sizeof(Sample) * static_cast<size_t>(stride) * pic->GetTotalHeight(c));
}

static void EnsureInit(xref_session *s) {
  if (s->inited) return;
  Qp base_qp(s->pic_qp, ChromaFormat::k420, s->bitdepth, s->lambda, s->chroma_table, s->off_u, s->off_v);
  s->pic_data->Init(s->segment, base_qp, false);
  s->inited = true;
}

// Build the leaf CUs of the picture from the shared descriptor array and mark them in the
// 4x4 CU map (PictureData::MarkUsedInPic), as the encoder does after deciding a CTU.
void xref_session_set_cus(xref_session *s, const xvcb200_cu *cus, int n) {
  EnsureInit(s);
  s->cus.clear();
  for (int i = 0; i < n; i++) {
    const xvcb200_cu &d = cus[i];
    CodingUnit *cu = s->pic_data->CreateCu(CuTree::Primary, d.depth, d.x, d.y, d.w, d.h);
    cu->SetPredMode((d.flags & XVCB200_CU_INTRA) ? PredictionMode::kIntra : PredictionMode::kInter);
    if (d.flags & XVCB200_CU_INTRA) {      // DC: the mode-dependent coefficient scan stays diagonal (transform.cc:1614-1636)
      cu->SetIntraModeLuma(IntraMode::kDc);
      cu->SetIntraModeChroma(IntraChromaMode::kDmChroma);
    }
    cu->SetQp(d.qp);
    cu->SetFullpelMv((d.flags & XVCB200_CU_FULLPEL_MV) != 0);
    cu->SetCbf(YuvComponent::kY, (d.flags & XVCB200_CU_CBF_Y) != 0);
    cu->SetCbf(YuvComponent::kU, (d.flags & XVCB200_CU_CBF_U) != 0);
    cu->SetCbf(YuvComponent::kV, (d.flags & XVCB200_CU_CBF_V) != 0);
    const bool l0 = d.ref_idx[0] >= 0, l1 = d.ref_idx[1] >= 0;
    cu->SetInterDir(l0 && l1 ? InterDir::kBi : (l1 ? InterDir::kL1 : InterDir::kL0));
    for (int l = 0; l < 2; l++) {
      cu->SetRefIdx(d.ref_idx[l], static_cast<RefPicList>(l));
      cu->SetMv(MotionVector(d.mv[l][0], d.mv[l][1]), static_cast<RefPicList>(l));
    }
    cu->SetTransformType(YuvComponent::kY, TransformType::kDefault, TransformType::kDefault);
    cu->SetTransformType(YuvComponent::kU, TransformType::kDefault, TransformType::kDefault);
    s->pic_data->MarkUsedInPic(cu);
    s->cus.push_back(cu);
  }
}

// Reads back what the pipeline changed in the CUs (cbf, chosen mv/ref).
void xref_session_get_cus(xref_session *s, xvcb200_cu *cus, int n) {
  for (int i = 0; i < n; i++) {
    const CodingUnit *cu = s->cus[i];
    xvcb200_cu &d = cus[i];
    d.flags &= ~(XVCB200_CU_CBF_Y | XVCB200_CU_CBF_U | XVCB200_CU_CBF_V);
    if (cu->GetCbf(YuvComponent::kY)) d.flags |= XVCB200_CU_CBF_Y;
    if (cu->GetCbf(YuvComponent::kU)) d.flags |= XVCB200_CU_CBF_U;
    if (cu->GetCbf(YuvComponent::kV)) d.flags |= XVCB200_CU_CBF_V;
    for (int l = 0; l < 2; l++) {
      d.ref_idx[l] = static_cast<int8_t>(cu->GetRefIdx(static_cast<RefPicList>(l)));
      const MotionVector &mv = cu->GetMv(static_cast<RefPicList>(l), MvCorner::kDefault);
      d.mv[l][0] = mv.x; d.mv[l][1] = mv.y;
    }
  }
}

// Transform modes of the session's CUs (xvcb200_set_tu_modes) + the luma intra mode the reference derives the
// coefficient scan from (TransformHelper::DetermineScanOrder; chroma follows luma: kDmChroma).  The scan
// entries of `modes` are NOT used here -- the reference computes its own from the intra mode.
void xref_session_set_tu_modes(xref_session *s, const xvcb200_tu_mode *modes, const uint8_t *intra_luma_mode, int n) {
  for (int i = 0; i < n && i < static_cast<int>(s->cus.size()); i++) {
    CodingUnit *cu = s->cus[i];
    if (cu->IsIntra() && intra_luma_mode) {
      cu->SetIntraModeLuma(static_cast<IntraMode>(intra_luma_mode[i]));
      cu->SetIntraModeChroma(IntraChromaMode::kDmChroma);
    }
    if (!modes) continue;
    cu->SetTransformType(YuvComponent::kY, static_cast<TransformType>(modes[i].tx_ver), static_cast<TransformType>(modes[i].tx_hor));
    for (int c = 0; c < 3; c++) {
      YuvComponent comp = static_cast<YuvComponent>(c);
      cu->SetTransformSkip(comp, cu->CanTransformSkip(comp) && ((modes[i].tskip >> c) & 1));
    }
  }
}

// TransformHelper::DetermineScanOrder for every CU and component (0 diagonal, 1 horizontal, 2 vertical)
void xref_session_scan_orders(xref_session *s, uint8_t *out /* n x 3 */) {
  for (size_t i = 0; i < s->cus.size(); i++)
    for (int c = 0; c < 3; c++) {
      const ScanOrder o = TransformHelper::DetermineScanOrder(*s->cus[i], static_cast<YuvComponent>(c));
      out[3 * i + c] = o == ScanOrder::kDiagonal ? 0 : (o == ScanOrder::kHorizontal ? 1 : 2);
    }
}

// ---------------------------------------------------------------- motion estimation
namespace {
struct MeWorker {
  explicit MeWorker(xref_session *s)
      : search(Simd(s->use_simd, s->bitdepth), *s->pic_data, *s->orig, *s->rec,
               *s->pic_data->GetRefPicLists(), s->settings) {}
  InterSearch search;
};

// One InterSearch::MotionEstNormal call (TZ + sub-pel), uni-prediction, TOrig = Sample.
void RunMeJob(xref_session *s, MeWorker *w, const xvcb200_me_job &job, double lambda, xvcb200_me_result *out) {
  CodingUnit *cu = s->cus[job.cu];
  const int list = job.list;
  const int ref_idx = job.ref_slot;   // in the shim, ref_slot indexes refs[list]
  const YuvPicture *ref_pic = s->refs[list][ref_idx].get();
  Qp qp(cu->GetQp(YuvComponent::kY), ChromaFormat::k420, s->bitdepth, lambda, s->chroma_table, s->off_u, s->off_v);
  MotionVector mvp(job.mvp[0], job.mvp[1]);
  MvFullpel clip_min, clip_max;
  w->search.DetermineMinMaxMv(*cu, *ref_pic, mvp, job.search_range, &clip_min, &clip_max);
  SampleMetric fullpel_metric(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth,
                              w->search.GetFullpelMetric(*cu));
  TzSearch tz(*s->orig, w->search, s->settings, job.search_range);
  MvFullpel prev(job.prev[0], job.prev[1]);
  // expose cost_best through a second pass is not possible; recompute below
  MvFullpel mv_full = tz.Search(*cu, qp, fullpel_metric, mvp, *ref_pic, clip_min, clip_max, prev);
  out->mv_fullpel[0] = mv_full.x; out->mv_fullpel[1] = mv_full.y;
  {
    // cost of the winner = dist + ((lambda*bits)>>16), as CheckCostBest computed it
    const Sample *o = s->orig->GetSamplePtr(YuvComponent::kY, cu->GetPosX(YuvComponent::kY), cu->GetPosY(YuvComponent::kY));
    const Sample *r = ref_pic->GetSamplePtr(YuvComponent::kY, cu->GetPosX(YuvComponent::kY) + mv_full.x,
                                            cu->GetPosY(YuvComponent::kY) + mv_full.y);
    Distortion d = fullpel_metric.CompareSample(qp, YuvComponent::kY, cu->GetWidth(YuvComponent::kY),
                                                cu->GetHeight(YuvComponent::kY), o, s->orig->GetStride(YuvComponent::kY),
                                                r, ref_pic->GetStride(YuvComponent::kY));
    uint32_t lam = static_cast<uint32_t>(std::floor(65536.0 * qp.GetLambdaSqrt()));
    Bits bits = InterSearch::GetMvdBitsFullpel(mvp, mv_full.x, mv_full.y, cu->GetFullpelMv() ? MvDelta::kPrecisionShift : 0);
    out->cost_fullpel = static_cast<uint32_t>(d + ((lam * bits) >> 16));
  }
  SampleMetric subpel_metric(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth,
                             w->search.GetSubpelMetric(*cu));
  SampleBufferStorage pred(constants::kMaxBlockSize, constants::kMaxBlockSize);
  auto orig_buffer = s->orig->GetSampleBuffer(YuvComponent::kY, cu->GetPosX(YuvComponent::kY), cu->GetPosY(YuvComponent::kY));
  Distortion dist = std::numeric_limits<Distortion>::max();
  MotionVector mv;
  if (cu->GetFullpelMv()) {
    mv = MotionVector(mv_full);
    dist = w->search.GetSubpelDist(*cu, qp, *ref_pic, subpel_metric, mv, orig_buffer, &pred);
    out->cost = static_cast<uint32_t>(dist);
  } else {
    mv = w->search.SubpelSearch(*cu, qp, subpel_metric, *ref_pic, mvp, mv_full, orig_buffer, &pred, &dist);
    uint32_t lam = static_cast<uint32_t>(std::floor(65536.0 * qp.GetLambdaSqrt()));
    out->cost = static_cast<uint32_t>(dist + ((lam * InterSearch::GetMvdBits(mvp, mv, 0)) >> 16));
  }
  out->mv[0] = mv.x; out->mv[1] = mv.y;
  out->dist = static_cast<uint32_t>(dist);
  out->num_sad = 0;
}
}  // namespace

// jobs[i].ref_slot = ref_idx within jobs[i].list.  lambda is Qp lambda (not sqrt).
void xref_me_search(xref_session *s, const xvcb200_me_job *jobs, int n, double lambda, int threads,
                    xvcb200_me_result *results) {
  EnsureInit(s);
  ParallelFor<MeWorker>(n, threads, [s]() { return new MeWorker(s); },
                        [&](MeWorker *w, int i) { RunMeJob(s, w, jobs[i], lambda, &results[i]); });
}

// TzSearch::Search alone (no sub-pel), for parity of the integer search.
void xref_tz_search(xref_session *s, const xvcb200_me_job *jobs, int n, double lambda, int32_t *mv_out) {
  EnsureInit(s);
  MeWorker w(s);
  for (int i = 0; i < n; i++) {
    const xvcb200_me_job &job = jobs[i];
    CodingUnit *cu = s->cus[job.cu];
    const YuvPicture *ref_pic = s->refs[job.list][job.ref_slot].get();
    Qp qp(cu->GetQp(YuvComponent::kY), ChromaFormat::k420, s->bitdepth, lambda, s->chroma_table, s->off_u, s->off_v);
    MotionVector mvp(job.mvp[0], job.mvp[1]);
    MvFullpel clip_min, clip_max;
    w.search.DetermineMinMaxMv(*cu, *ref_pic, mvp, job.search_range, &clip_min, &clip_max);
    SampleMetric metric(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth, w.search.GetFullpelMetric(*cu));
    TzSearch tz(*s->orig, w.search, s->settings, job.search_range);
    MvFullpel mv = tz.Search(*cu, qp, metric, mvp, *ref_pic, clip_min, clip_max, MvFullpel(job.prev[0], job.prev[1]));
    mv_out[2 * i] = mv.x; mv_out[2 * i + 1] = mv.y;
  }
}

// InterSearch::FullSearch on the weighted original 2*orig - other_pred (bi-pred refinement).
// other_pred = s->pred (set via xref_session_set_pred).  ref taken from jobs[i].ref_slot in
// list jobs[i].other_pred_slot (reused as "list" in the shim).
void xref_full_search(xref_session *s, const xvcb200_fullsearch_job *jobs, int n, double lambda,
                      xvcb200_me_result *results) {
  EnsureInit(s);
  MeWorker w(s);
  for (int i = 0; i < n; i++) {
    const xvcb200_fullsearch_job &job = jobs[i];
    CodingUnit *cu = s->cus[job.cu];
    const int list = job.other_pred_slot;
    const YuvPicture *ref_pic = s->refs[list][job.ref_slot].get();
    const int x = cu->GetPosX(YuvComponent::kY), y = cu->GetPosY(YuvComponent::kY);
    const int cw = cu->GetWidth(YuvComponent::kY), ch = cu->GetHeight(YuvComponent::kY);
    Qp qp(cu->GetQp(YuvComponent::kY), ChromaFormat::k420, s->bitdepth, lambda, s->chroma_table, s->off_u, s->off_v);
    w.search.bipred_orig_buffer_.SubtractWeighted(cw, ch, s->orig->GetSampleBuffer(YuvComponent::kY, x, y),
                                                   SampleBufferConst(s->pred->GetSamplePtr(YuvComponent::kY, x, y),
                                                                     s->pred->GetStride(YuvComponent::kY)));
    MotionVector mvp(job.mvp[0], job.mvp[1]);
    MotionVector center(job.center[0], job.center[1]);
    MvFullpel clip_min, clip_max;
    w.search.DetermineMinMaxMv(*cu, *ref_pic, center, job.range, &clip_min, &clip_max);
    SampleMetric metric(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth, w.search.GetFullpelMetric(*cu));
    MvFullpel mv = w.search.FullSearch(*cu, qp, metric, mvp, *ref_pic, clip_min, clip_max);
    std::memset(&results[i], 0, sizeof(results[i]));
    results[i].mv_fullpel[0] = mv.x; results[i].mv_fullpel[1] = mv.y;
  }
}

// ---------------------------------------------------------------- motion compensation
// InterPrediction::MotionCompensation for every CU and component into s->pred.
void xref_motion_compensate(xref_session *s, int threads) {
  EnsureInit(s);
  const int n = static_cast<int>(s->cus.size());
  const auto &simd = Simd(s->use_simd, s->bitdepth);
  ParallelFor<InterPrediction>(
      n, threads, [&]() { return new InterPrediction(simd.inter_prediction, *s->rec, s->bitdepth); },
      [&](InterPrediction *ip, int i) {
        CodingUnit *cu = s->cus[i];
        if (cu->IsIntra()) return;
        for (int c = 0; c < 3; c++) {
          YuvComponent comp = static_cast<YuvComponent>(c);
          SampleBuffer pb = s->pred->GetSampleBuffer(comp, cu->GetPosX(comp), cu->GetPosY(comp));
          ip->MotionCompensation(*cu, comp, &pb);
        }
      });
}

// Affine CUs: CodingUnit::SetUseAffine + SetMv(MotionVector3) (coding_unit.h:250-257, 287), then the
// reference's own InterPrediction::MotionCompensation, which reaches MotionCompAffine through
// MotionCompRefList (inter_prediction.cc:1021-1023).  The CUs are restored to translational after.
void xref_motion_compensate_affine(xref_session *s, const xvcb200_affine_cu *aff, int n, int threads) {
  EnsureInit(s);
  const auto &simd = Simd(s->use_simd, s->bitdepth);
  ParallelFor<InterPrediction>(
      n, threads, [&]() { return new InterPrediction(simd.inter_prediction, *s->rec, s->bitdepth); },
      [&](InterPrediction *ip, int i) {
        CodingUnit *cu = s->cus[aff[i].cu];
        if (cu->IsIntra()) return;
        MotionVector keep[2];
        for (int l = 0; l < 2; l++) {
          const RefPicList list = static_cast<RefPicList>(l);
          keep[l] = cu->GetMv(list, MvCorner::kDefault);
          MotionVector3 mv3;
          for (int k = 0; k < 3; k++) mv3[k] = MotionVector(aff[i].mv[l][k][0], aff[i].mv[l][k][1]);
          cu->SetMv(mv3, list);
        }
        cu->SetUseAffine(true);
        for (int c = 0; c < 3; c++) {
          YuvComponent comp = static_cast<YuvComponent>(c);
          SampleBuffer pb = s->pred->GetSampleBuffer(comp, cu->GetPosX(comp), cu->GetPosY(comp));
          ip->MotionCompensation(*cu, comp, &pb);
        }
        cu->SetUseAffine(false);
        for (int l = 0; l < 2; l++) cu->SetMv(keep[l], static_cast<RefPicList>(l));
      });
}

// LIC CUs: CodingUnit::SetUseLic(true), then the reference's own MotionCompensation
// (LocalIlluminationComp / DeriveLicParams read s->rec, the InterPrediction's rec_pic_, around the CU).
// Serial: DeriveLicParams looks at neighbour CUs, which no other thread may be modifying.
void xref_motion_compensate_lic(xref_session *s, const xvcb200_lic_cu *lic, int n) {
  EnsureInit(s);
  const auto &simd = Simd(s->use_simd, s->bitdepth);
  std::thread([&]() {            // fresh thread: unrestricted thread_local Restrictions (see ParallelFor)
    InterPrediction ip(simd.inter_prediction, *s->rec, s->bitdepth);
    for (int i = 0; i < n; i++) {
      CodingUnit *cu = s->cus[lic[i].cu];
      if (cu->IsIntra()) continue;
      cu->SetUseLic(true);
      for (int c = 0; c < 3; c++) {
        YuvComponent comp = static_cast<YuvComponent>(c);
        SampleBuffer pb = s->pred->GetSampleBuffer(comp, cu->GetPosX(comp), cu->GetPosY(comp));
        ip.MotionCompensation(*cu, comp, &pb);
      }
      cu->SetUseLic(false);
    }
  }).join();
}

// What the reference sees as the CU above / left of every CU (CodingUnit::GetCodingUnitAbove / Left):
// fills above_x/above_y/left_x/left_y of lic[i] for lic[i].cu -- checks the host-side CU map of the tests.
void xref_lic_neighbours(xref_session *s, xvcb200_lic_cu *lic, int n) {
  EnsureInit(s);
  for (int i = 0; i < n; i++) {
    const CodingUnit *cu = s->cus[lic[i].cu];
    const CodingUnit *a = cu->GetCodingUnitAbove(), *l = cu->GetCodingUnitLeft();
    lic[i].above_x = a ? static_cast<int16_t>(a->GetPosX(YuvComponent::kY)) : -1;
    lic[i].above_y = a ? static_cast<int16_t>(a->GetPosY(YuvComponent::kY)) : -1;
    lic[i].left_x = l ? static_cast<int16_t>(l->GetPosX(YuvComponent::kY)) : -1;
    lic[i].left_y = l ? static_cast<int16_t>(l->GetPosY(YuvComponent::kY)) : -1;
  }
}

// ---------------------------------------------------------------- T/Q/recon chain
// The body of TransformEncoder::TransformAndReconstruct (transform_encoder.cc:203-285),
// driven class by class with RdoQuant::QuantFast in place of QuantRdo (rdo_quant is a
// compile-time constant in the reference; QuantFast is its non-RDO quantiser).
void xref_tq_reconstruct(xref_session *s, int threads, xvcb200_tu_result *results) {
  EnsureInit(s);
  const int n = static_cast<int>(s->cus.size());
  const int bd = s->bitdepth;
  struct TqState {
    TqState(xref_session *s, int bd)
        : fwd(bd), inv(bd), q(bd, s->settings),
          ssd(Simd(s->use_simd, bd).sample_metric, bd, MetricType::kSsd),
          resi_orig(64, 64), resi(64, 64), tmp(64, 64), lev(64, 64) {}
    ForwardTransform fwd;
    InverseTransform inv;
    Quantize dq;
    RdoQuant q;
    SampleMetric ssd;
    ResidualBufferStorage resi_orig, resi;
    CoeffBufferStorage tmp, lev;
  };
  const Sample max_pel = static_cast<Sample>((1 << bd) - 1);
  ParallelFor<TqState>(n, threads, [&]() { return new TqState(s, bd); }, [&](TqState *t, int i) {
    CodingUnit *cu = s->cus[i];
    const Qp &qp = cu->GetQp();
    Qp luma_w(cu->GetQp(YuvComponent::kY), ChromaFormat::k420, bd, 1.0, 0, 0, 0);  // weight 1.0
    for (int c = 0; c < 3; c++) {
      YuvComponent comp = static_cast<YuvComponent>(c);
      const int x = cu->GetPosX(comp), y = cu->GetPosY(comp), w = cu->GetWidth(comp), h = cu->GetHeight(comp);
      SampleBufferConst orig = s->orig->GetSampleBuffer(comp, x, y);
      SampleBuffer pred = s->pred->GetSampleBuffer(comp, x, y);
      SampleBuffer reco = s->rec->GetSampleBuffer(comp, x, y);
      t->resi_orig.Subtract(w, h, orig, pred);
      const bool skip_transform = cu->GetTransformSkip(comp);      // transform_encoder.cc:213-227
      if (!skip_transform) t->fwd.Transform(*cu, comp, t->resi_orig, &t->tmp);
      else t->fwd.TransformSkip(w, h, t->resi_orig, &t->tmp);
      int nz = t->q.QuantFast(*cu, comp, qp, cu->GetPicType(), t->tmp.GetDataPtr(), t->tmp.GetStride(),
                              t->lev.GetDataPtr(), t->lev.GetStride());
      cu->SetDcCoeffOnly(comp, false);   // the batched path never takes the DC shortcut
      const bool cbf = nz != 0;
      cu->SetCbf(comp, cbf);
      const int pw = s->orig->GetWidth(comp);
      for (int yy = 0; yy < h; yy++)
        for (int xx = 0; xx < w; xx++)
          s->coeff[c][static_cast<size_t>(y + yy) * pw + x + xx] =
              cbf ? t->lev.GetDataPtr()[yy * t->lev.GetStride() + xx] : 0;
      if (cbf) {
        t->dq.Inverse(comp, qp, w, h, bd, t->lev.GetDataPtr(), t->lev.GetStride(), t->tmp.GetDataPtr(),
                      t->tmp.GetStride());
        if (!skip_transform) t->inv.Transform(*cu, comp, t->tmp, &t->resi);
        else t->inv.TransformSkip(w, h, t->tmp, &t->resi);
        reco.AddClip(w, h, pred, t->resi, 0, max_pel);
      } else {
        reco.CopyFrom(w, h, pred);
      }
      if (results) {
        results[3 * i + c].num_non_zero = nz;
        results[3 * i + c].ssd = static_cast<uint32_t>(
            t->ssd.Compare(luma_w, YuvComponent::kY, w, h, orig.GetDataPtr(), orig.GetStride(),
                           reco.GetDataPtr(), reco.GetStride()));
      }
    }
  });
}

// ---------------------------------------------------------------- in-loop filter + padding
void xref_deblock_picture(xref_session *s, int beta_offset, int tc_offset) {
  EnsureInit(s);
  DeblockingFilter f(s->pic_data.get(), s->rec.get(), beta_offset, tc_offset);
  f.DeblockPicture();
}

// DeblockPicture with the two things the CU descriptors alone do not carry: affine CUs (corner vectors,
// deblocking_filter.cc:166-176) and the secondary CU tree of an intra picture (:65-68, 88-91).  The affine
// CUs keep SetUseAffine / SetMv(MotionVector3) for the duration of the call.
void xref_deblock_picture_ext(xref_session *s, int beta_offset, int tc_offset, const xvcb200_affine_cu *aff, int n_aff,
                              const xvcb200_cu *chroma_cus, int n_chroma) {
  EnsureInit(s);
  for (int i = 0; i < n_aff; i++) {
    CodingUnit *cu = s->cus[aff[i].cu];
    cu->SetUseAffine(true);
    for (int l = 0; l < 2; l++) {
      if (cu->GetRefIdx(static_cast<RefPicList>(l)) < 0) continue;
      MotionVector3 mv;
      for (int k = 0; k < 3; k++) mv[k] = MotionVector(aff[i].mv[l][k][0], aff[i].mv[l][k][1]);
      cu->SetMv(mv, static_cast<RefPicList>(l));
    }
  }
  if (n_chroma > 0 && s->pic_data->HasSecondaryCuTree())
    for (int i = 0; i < n_chroma; i++) {
      const xvcb200_cu &d = chroma_cus[i];
      CodingUnit *cu = s->pic_data->CreateCu(CuTree::Secondary, d.depth, d.x, d.y, d.w, d.h);
      cu->SetPredMode(PredictionMode::kIntra);
      cu->SetQp(d.qp);
      s->pic_data->MarkUsedInPic(cu);
    }
  DeblockingFilter f(s->pic_data.get(), s->rec.get(), beta_offset, tc_offset);
  f.DeblockPicture();
  for (int i = 0; i < n_aff; i++) {
    CodingUnit *cu = s->cus[aff[i].cu];
    cu->SetUseAffine(false);
    for (int l = 0; l < 2; l++)
      if (cu->GetRefIdx(static_cast<RefPicList>(l)) >= 0)
        cu->SetMv(MotionVector(aff[i].mv[l][0][0], aff[i].mv[l][0][1]), static_cast<RefPicList>(l));
  }
}
int xref_session_has_secondary_tree(xref_session *s) { EnsureInit(s); return s->pic_data->HasSecondaryCuTree() ? 1 : 0; }

void xref_pad_border_rec(xref_session *s) { s->rec->PadBorder(); }

void xref_session_get_rec_padded(xref_session *s, int comp, uint16_t *out) {
  const YuvPicture *pic = s->rec.get();
  YuvComponent c = static_cast<YuvComponent>(comp);
  int stride = static_cast<int>(pic->GetStride(c));
  int off_x = (stride - pic->GetWidth(c)) / 2;
  int off_y = (pic->GetTotalHeight(c) - pic->GetHeight(c)) / 2;
  std::memcpy(out, pic->GetSamplePtr(c, -off_x, -off_y),
              sizeof(Sample) * static_cast<size_t>(stride) * pic->GetTotalHeight(c));
}

// ---------------------------------------------------------------- whole-picture hot path
// Same step as xvcb200_encode_picture, executed by the reference's own classes:
// InterSearch::SearchMotion's control flow (inter_search.cc:199-259: SearchRefIdx per list :456-578,
// SearchBiIterative :392-433, final choice :245-258) around the reference's TzSearch::Search /
// FullSearch / SubpelSearch / MotionCompensation / SubtractWeighted / GetInterPredBits, with the two
// simplifications of the batched design (include/xvc_b200.h, xvcb200_picture_params): one predictor
// per list (the CU's mv[list] as handed over) and previous_fullpel_ = 0 -> MC -> T/Q/recon ->
// deblock -> pad.  Every sample, distortion and bit count is produced by reference code.
namespace {
struct PipeWorker {
  explicit PipeWorker(xref_session *s)
      : me(s), base_qp(s->pic_qp, ChromaFormat::k420, s->bitdepth, s->lambda, s->chroma_table, s->off_u, s->off_v),
        writer(base_qp, s->pic_data->GetPredictionType(), &bw), pred(constants::kMaxBlockSize, constants::kMaxBlockSize) {}
  MeWorker me;
  Qp base_qp;
  BitWriter bw;
  SyntaxWriter writer;
  SampleBufferStorage pred;
};

void SearchMotionCu(xref_session *s, PipeWorker *w, const xvcb200_picture_params *p, const int R[2], double lambda, int i,
                    const xvcb200_cu &in, const int32_t *mvp_cols /* J x 2 or null */, xvcb200_me_result *res /* J entries */) {
  CodingUnit *cu = s->cus[i];
  if (in.flags & (XVCB200_CU_INTRA | XVCB200_CU_SKIP_ME)) return;
  InterSearch &search = w->me.search;
  const YuvComponent comp = YuvComponent::kY;
  Qp qp(cu->GetQp(comp), ChromaFormat::k420, s->bitdepth, lambda, s->chroma_table, s->off_u, s->off_v);
  const uint32_t lam = static_cast<uint32_t>(std::floor(65536.0 * qp.GetLambdaSqrt()));
  // predictor of (list, ref_idx): the caller's column (GetMvpList is per ref_idx), else the CU's mv[list]
  auto mvp_of = [&](int l, int r) {
    const int col = l * R[0] + r;
    return mvp_cols ? MotionVector(mvp_cols[2 * col], mvp_cols[2 * col + 1]) : MotionVector(in.mv[l][0], in.mv[l][1]);
  };
  const Distortion kMax = std::numeric_limits<Distortion>::max();
  auto list_of = [](int l) { return static_cast<RefPicList>(l); };
  auto dir_of = [](int l) { return l == 0 ? InterDir::kL0 : InterDir::kL1; };
  int uni_ref[2] = {-1, -1}, l1u_ref = -1, bi_ref[2] = {-1, -1};
  MotionVector uni_mv[2], l1u_mv, bi_mv[2];
  Distortion cost_uni[2] = {kMax, kMax}, cost_l1u = kMax, cost_bi = kMax;
  for (int l = 0; l < 2; l++) {
    cu->SetInterDir(dir_of(l));
    cu->SetMv(MotionVector(), list_of(1 - l));          // SearchRefIdx clears the other list (:476-480)
    cu->SetRefIdx(-1, list_of(1 - l));
    for (int r = 0; r < R[l]; r++) {
      const int col = l * R[0] + r;
      const int dup = (l == 1 && p->bits_mode) ? search.same_poc_in_l0_mapping_[r] : -1;
      if (dup >= 0) {
        res[col] = res[dup];                              // :536-543
      } else {
        xvcb200_me_job job;
        job.cu = i; job.ref_slot = r; job.list = l; job.search_range = p->search_range[l][r];
        job.mvp[0] = mvp_of(l, r).x; job.mvp[1] = mvp_of(l, r).y; job.prev[0] = job.prev[1] = 0;
        RunMeJob(s, &w->me, job, lambda, &res[col]);
      }
      const MotionVector mv(res[col].mv[0], res[col].mv[1]);
      Distortion cost = res[col].cost;
      if (p->bits_mode) {
        cu->SetRefIdx(r, list_of(l));
        cu->SetMvpIdx(0, list_of(l));
        cu->SetMv(mv, list_of(l));
        search.SetMvd(cu, list_of(l), mvp_of(l, r), mv);
        cost = res[col].dist + ((search.GetInterPredBits(*cu, w->writer) * lam) >> 16);
      }
      if (cost < cost_uni[l]) { cost_uni[l] = cost; uni_ref[l] = r; uni_mv[l] = mv; }
      if (l == 1 && dup < 0 && cost < cost_l1u) { cost_l1u = cost; l1u_ref = r; l1u_mv = mv; }
    }
  }
  if (p->bi_iterations > 0 && uni_ref[0] >= 0 && uni_ref[1] >= 0) {
    const int x = cu->GetPosX(comp), y = cu->GetPosY(comp), cw = cu->GetWidth(comp), ch = cu->GetHeight(comp);
    SampleBufferConst orig_luma = s->orig->GetSampleBuffer(comp, x, y);
    SampleMetric fullpel_metric(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth, search.GetFullpelMetric(*cu));
    SampleMetric subpel_metric(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth, search.GetSubpelMetric(*cu));
    for (int l = 0; l < 2; l++) { bi_ref[l] = uni_ref[l]; bi_mv[l] = uni_mv[l]; }
    int sl = cost_uni[0] <= cost_uni[1] ? 1 : 0;         // the list with the higher uni cost first (:402-403)
    for (int it = 0; it < p->bi_iterations; it++) {
      const int other = 1 - sl;
      cu->SetInterDir(dir_of(other));
      cu->SetRefIdx(bi_ref[other], list_of(other));
      cu->SetMv(bi_mv[other], list_of(other));
      search.MotionCompensation(*cu, comp, &search.bipred_pred_buffer_);
      search.bipred_orig_buffer_.SubtractWeighted(cw, ch, orig_luma, search.bipred_pred_buffer_);
      cu->SetInterDir(InterDir::kBi);
      const Distortion prev_best = cost_bi;
      for (int r = 0; r < R[sl]; r++) {
        const int col = sl * R[0] + r;
        const YuvPicture *ref_pic = s->refs[sl][r].get();
        const MotionVector bootstrap(res[col].mv[0], res[col].mv[1]);   // GetBestUniPredMv (:497)
        MvFullpel clip_min, clip_max;
        search.DetermineMinMaxMv(*cu, *ref_pic, bootstrap, EncoderSettings::inter_search_range_bi, &clip_min, &clip_max);
        const MotionVector mvp_sl = mvp_of(sl, r);
        const MvFullpel mv_full = search.FullSearch(*cu, qp, fullpel_metric, mvp_sl, *ref_pic, clip_min, clip_max);
        Distortion dist = kMax;
        MotionVector mv;
        if (cu->GetFullpelMv()) {
          mv = MotionVector(mv_full);
          dist = search.GetSubpelDist(*cu, qp, *ref_pic, subpel_metric, mv, search.bipred_orig_buffer_, &w->pred);
        } else {
          mv = search.SubpelSearch(*cu, qp, subpel_metric, *ref_pic, mvp_sl, mv_full, search.bipred_orig_buffer_, &w->pred, &dist);
        }
        dist >>= 1;                                        // MotionEstNormal, :660
        if (p->bi_iterations > 1) { res[col].mv[0] = mv.x; res[col].mv[1] = mv.y; }   // SetBestUniPredMv (:549-553)
        cu->SetRefIdx(r, list_of(sl));
        cu->SetMvpIdx(0, list_of(sl));
        cu->SetMvpIdx(0, list_of(other));
        cu->SetMv(mv, list_of(sl));
        search.SetMvd(cu, list_of(sl), mvp_sl, mv);
        search.SetMvd(cu, list_of(other), mvp_of(other, bi_ref[other]), bi_mv[other]);
        const Distortion cost = dist + ((search.GetInterPredBits(*cu, w->writer) * lam) >> 16);
        if (cost < cost_bi) { cost_bi = cost; bi_ref[sl] = r; bi_mv[sl] = mv; }
      }
      if (cost_bi == prev_best) break;                    // :425-427
      sl = other;
    }
  }
  int ref[2] = {-1, -1};
  MotionVector mv[2];
  if (p->bi_iterations > 0 && cost_bi != kMax && cost_bi <= cost_uni[0] && cost_bi <= cost_l1u) {
    for (int l = 0; l < 2; l++) { ref[l] = bi_ref[l]; mv[l] = bi_mv[l]; }
  } else if (cost_uni[0] <= cost_l1u) {
    ref[0] = uni_ref[0]; mv[0] = uni_mv[0];
  } else {
    ref[1] = l1u_ref; mv[1] = l1u_mv;
  }
  for (int l = 0; l < 2; l++) {
    cu->SetRefIdx(ref[l], list_of(l));
    cu->SetMv(mv[l], list_of(l));
  }
  cu->SetInterDir(ref[0] >= 0 && ref[1] >= 0 ? InterDir::kBi : (ref[1] >= 0 ? InterDir::kL1 : InterDir::kL0));
}
}  // namespace

void xref_encode_picture_mvp(xref_session *s, const xvcb200_picture_params *p, const xvcb200_cu *cus_in, int n,
                             const int32_t *mvp /* n x J x 2 or null: xvcb200_set_mv_predictors */, int threads,
                             xvcb200_me_result *me_results, xvcb200_tu_result *tu_results, xvcb200_cu *cus_out) {
  xref_session_set_cus(s, cus_in, n);
  s->pic_data->force_bipred_l1_mvd_zero_ = false;       // the low-delay "L1 mvd = 0" rule is not part of the batched step
  s->settings.fast_inter_pred_bits = p->bits_mode ? 1 : 0;
  const double lambda = p->lambda_sqrt * p->lambda_sqrt;
  const int R[2] = {p->num_ref[0] > 0 ? p->num_ref[0] : 1, p->pic_type == 1 ? 0 : (p->num_ref[1] > 0 ? p->num_ref[1] : 1)};
  const int J = R[0] + R[1];
  std::vector<xvcb200_me_result> res(static_cast<size_t>(n) * J);
  std::memset(res.data(), 0, res.size() * sizeof(res[0]));
  // lambda_sqrt is authoritative: build Qp from lambda = lambda_sqrt^2 and check the sqrt
  // round-trips (it does for the values bench.py uses; asserted in tests).
  ParallelFor<PipeWorker>(n, threads, [s]() { return new PipeWorker(s); },
                          [&](PipeWorker *w, int i) {
                            SearchMotionCu(s, w, p, R, lambda, i, cus_in[i], mvp ? mvp + static_cast<size_t>(i) * J * 2 : nullptr,
                                           &res[static_cast<size_t>(i) * J]);
                          });
  if (me_results) std::memcpy(me_results, res.data(), res.size() * sizeof(res[0]));
  xref_motion_compensate(s, threads);
  xref_tq_reconstruct(s, threads, tu_results);
  if (p->deblock) xref_deblock_picture(s, p->beta_offset, p->tc_offset);
  if (p->pad) s->rec->PadBorder();
  if (cus_out) {
    std::memcpy(cus_out, cus_in, sizeof(xvcb200_cu) * n);
    xref_session_get_cus(s, cus_out, n);
  }
}

void xref_encode_picture(xref_session *s, const xvcb200_picture_params *p, const xvcb200_cu *cus_in, int n,
                         int threads, xvcb200_me_result *me_results, xvcb200_tu_result *tu_results,
                         xvcb200_cu *cus_out) {
  xref_encode_picture_mvp(s, p, cus_in, n, nullptr, threads, me_results, tu_results, cus_out);
}

// Pins the control flow above to the reference's own InterSearch::SearchMotion: for a picture holding
// ONE CU (no neighbours: both predictors of GetMvpList are zero, previous_fullpel_ is zero) the two
// coincide when the CU's predictor is zero.  out[0] = SearchMotionCu, out[1] = InterSearch::SearchMotion
// (ref_idx / mv of the CU); costs[0..1] likewise.
void xref_search_motion_single(xref_session *s, const xvcb200_picture_params *p, const xvcb200_cu *cu_in, xvcb200_cu *out,
                               uint64_t *costs) {
  xvcb200_cu in = *cu_in;
  std::memset(in.mv, 0, sizeof(in.mv));
  xref_session_set_cus(s, &in, 1);
  s->pic_data->force_bipred_l1_mvd_zero_ = false;
  s->settings.fast_inter_pred_bits = 1;
  s->settings.bipred_refinement_iterations = p->bi_iterations;
  const double lambda = p->lambda_sqrt * p->lambda_sqrt;
  const int R[2] = {p->num_ref[0] > 0 ? p->num_ref[0] : 1, p->pic_type == 1 ? 0 : (p->num_ref[1] > 0 ? p->num_ref[1] : 1)};
  std::vector<xvcb200_me_result> res(R[0] + R[1]);
  CodingUnit *cu = s->cus[0];
  {
    PipeWorker w(s);
    SearchMotionCu(s, &w, p, R, lambda, 0, in, nullptr, res.data());
    out[0] = in;
    xref_session_get_cus(s, &out[0], 1);
    costs[0] = 0;
  }
  {
    PipeWorker w(s);
    Qp qp(cu->GetQp(YuvComponent::kY), ChromaFormat::k420, s->bitdepth, lambda, s->chroma_table, s->off_u, s->off_v);
    const InterSearchFlags flags = (in.flags & XVCB200_CU_FULLPEL_MV) ? InterSearchFlags::kFullPelMv : InterSearchFlags(0);
    costs[1] = p->pic_type == 1 ? w.me.search.SearchMotion(cu, qp, w.writer, flags | InterSearchFlags::kUniPredOnly, &w.pred)
                                : w.me.search.SearchMotion(cu, qp, w.writer, flags, &w.pred);
    out[1] = in;
    xref_session_get_cus(s, &out[1], 1);
  }
}

// ---------------------------------------------------------------- intra prediction
// For the CUs of `cus` in coding order (each one sees only the CUs before it, like the encoder
// while it walks the picture): the neighbour availability, the reference samples as
// IntraPrediction::FillReferenceState leaves them (from the session's reconstruction picture),
// the prediction of every mode (planar, DC, 65 angular) and, for luma, its SATD against the
// session's original picture.  comp: 0 luma, 1/2 chroma.
//   ref_out / filt_out : n x 2 x 129 samples (entries the block does not use are zeroed)
//   pred_out           : for CU i, 67 blocks of (w x h) tight samples, CUs back to back (or null)
//   satd_out           : n x 67 (luma only, or null)
void xref_intra_scan(xref_session *s, const xvcb200_cu *cus, int n, int comp_i, xvcb200_intra_job *jobs_out,
                     uint16_t *ref_out, uint16_t *filt_out, uint16_t *pred_out, uint32_t *satd_out) {
  EnsureInit(s);
  const YuvComponent comp = static_cast<YuvComponent>(comp_i);
  IntraPrediction ip(s->bitdepth);
  SampleMetric satd(Simd(s->use_simd, s->bitdepth).sample_metric, s->bitdepth, MetricType::kSatd);
  Qp qp(s->pic_qp, ChromaFormat::k420, s->bitdepth, 1.0, 1, 0, 0);
  const ptrdiff_t rs = IntraPrediction::kRefSampleStride_;
  std::vector<Sample> tmp(64 * 64 * 2);
  size_t pred_off = 0;
  for (int i = 0; i < n; i++) {
    const xvcb200_cu &d = cus[i];
    CodingUnit *cu = s->pic_data->CreateCu(CuTree::Primary, d.depth, d.x, d.y, d.w, d.h);
    cu->SetPredMode(PredictionMode::kIntra);
    cu->SetQp(d.qp);
    const int w = cu->GetWidth(comp), h = cu->GetHeight(comp), x = cu->GetPosX(comp), y = cu->GetPosY(comp);
    xvcb200_intra_job &j = jobs_out[i];
    std::memset(&j, 0, sizeof(j));
    j.x = x; j.y = y; j.w = static_cast<uint8_t>(w); j.h = static_cast<uint8_t>(h);
    j.has_left = x > 0; j.has_above = y > 0; j.has_above_left = x > 0 && y > 0;          // DetermineNeighbors, :688-707
    j.below_left = x > 0 ? static_cast<uint8_t>(cu->GetCuSizeBelowLeft(comp)) : 0;
    j.above_right = y > 0 ? static_cast<uint8_t>(cu->GetCuSizeAboveRight(comp)) : 0;
    IntraPrediction::RefState state;
    state.ref_samples.fill(0);
    state.ref_filtered.fill(0);
    ip.FillReferenceState(*cu, comp, *s->rec, &state);
    for (int k = 0; k < 2 * rs; k++) {
      const bool used = k <= w + h || (k >= rs && k < rs + w + h);
      ref_out[static_cast<size_t>(i) * 2 * rs + k] = used ? state.ref_samples[k] : 0;
      filt_out[static_cast<size_t>(i) * 2 * rs + k] = (used && comp_i == 0) ? state.ref_filtered[k] : 0;
    }
    for (int mode = 0; mode < kNbrIntraModesExt; mode++) {
      // horizontal modes of non-square blocks write the transposed block first: 64-sample stride, 128 rows
      SampleBuffer out(tmp.data(), 64);
      ip.Predict(static_cast<IntraMode>(mode), *cu, comp, state, *s->rec, &out);
      if (pred_out) {
        for (int r = 0; r < h; r++) std::memcpy(pred_out + pred_off + static_cast<size_t>(r) * w, tmp.data() + r * 64, sizeof(Sample) * w);
        pred_off += static_cast<size_t>(w) * h;
      }
      if (satd_out && comp_i == 0)
        satd_out[static_cast<size_t>(i) * kNbrIntraModesExt + mode] = static_cast<uint32_t>(
            satd.Compare(qp, comp, w, h, s->orig->GetSamplePtr(comp, x, y), s->orig->GetStride(comp), tmp.data(), 64));
    }
    s->pic_data->MarkUsedInPic(cu);
    s->cus.push_back(cu);
  }
}

// IntraPrediction::Predict(kLmChroma) -> PredLmChroma (intra_prediction.cc:560-584) for n CUs: U then V on
// one IntraPrediction object (the second component reuses the reduced luma of the first, :574-580).
// Reads s->rec (the CU's reconstructed luma + the luma / chroma neighbours).  pred_u / pred_v: the
// blocks back to back, (w/2) x (h/2) each.  Fresh thread: unrestricted Restrictions (see ParallelFor).
void xref_intra_lm_chroma(xref_session *s, const xvcb200_cu *cus, int n, uint16_t *pred_u, uint16_t *pred_v) {
  EnsureInit(s);
  std::thread([&]() {
    IntraPrediction ip(s->bitdepth);
    std::vector<Sample> tmp(64 * 64);
    size_t off = 0;
    for (int i = 0; i < n; i++) {
      const xvcb200_cu &d = cus[i];
      CodingUnit *cu = s->pic_data->CreateCu(CuTree::Primary, d.depth, d.x, d.y, d.w, d.h);
      cu->SetPredMode(PredictionMode::kIntra);
      cu->SetQp(d.qp);
      IntraPrediction::RefState state;
      state.ref_samples.fill(0);
      state.ref_filtered.fill(0);
      const int w = d.w / 2, h = d.h / 2;
      for (int c = 1; c <= 2; c++) {
        SampleBuffer out(tmp.data(), 64);
        ip.Predict(IntraMode::kLmChroma, *cu, static_cast<YuvComponent>(c), state, *s->rec, &out);
        uint16_t *dst = (c == 1 ? pred_u : pred_v) + off;
        for (int r = 0; r < h; r++) std::memcpy(dst + static_cast<size_t>(r) * w, tmp.data() + r * 64, sizeof(Sample) * w);
      }
      off += static_cast<size_t>(w) * h;
      s->pic_data->MarkUsedInPic(cu);
      s->cus.push_back(cu);
    }
  }).join();
}

// ---------------------------------------------------------------- bitstream conformance
// A real xvc bitstream whose inter picture carries decisions and levels made OUTSIDE the reference
// (by the GPU path): the reference encoder (public API, low delay, one reference) codes the key
// picture and provides all high-level plumbing; for the inter picture its own CTU loop is
// replaced by "build the CU tree from the given partition, fill every CU from the given
// decisions, write it with the reference's CuWriter / SyntaxWriter", and the picture checksum is
// taken over the reconstruction that was handed in.  The reference DECODER (xvcdec) then
// verifies that checksum against what it decodes.
struct xref_conf {
  const xvc_encoder_api *api = nullptr;
  xvc_encoder_parameters *params = nullptr;
  xvc_encoder *handle = nullptr;
  Encoder *enc = nullptr;
  int width = 0, height = 0, bitdepth = 0;
  struct Nal { uint32_t type, poc; std::vector<uint8_t> bytes; };
  std::vector<Nal> nals;
};

static void ConfCollect(xref_conf *c, xvc_enc_nal_unit *units, int n) {
  for (int i = 0; i < n; i++) {
    xref_conf::Nal nal;
    nal.type = units[i].stats.nal_unit_type;
    nal.poc = units[i].stats.poc;
    nal.bytes.assign(units[i].bytes, units[i].bytes + units[i].size);
    c->nals.push_back(std::move(nal));
  }
}

static std::shared_ptr<PictureEncoder> ConfFind(xref_conf *c, int poc) {
  for (auto &pe : c->enc->pic_encoders_)
    if (static_cast<int>(pe->GetPoc()) == poc) return pe;
  return nullptr;
}

xref_conf *xref_conf_create(int width, int height, int bitdepth, int qp) {
  xref_conf *c = new xref_conf();
  c->api = xvc_encoder_api_get();
  c->params = c->api->parameters_create();
  c->api->parameters_set_default(c->params);
  c->params->width = width; c->params->height = height;
  c->params->chroma_format = XVC_ENC_CHROMA_FORMAT_420;
  c->params->input_bitdepth = bitdepth; c->params->internal_bitdepth = bitdepth;
  c->params->framerate = 30;
  c->params->sub_gop_length = 1; c->params->low_delay = 1; c->params->num_ref_pics = 1;
  c->params->qp = qp; c->params->threads = 0; c->params->speed_mode = 2;
  c->params->checksum_mode = 1;               // a checksum for every picture
  c->handle = c->api->encoder_create(c->params);
  if (!c->handle) { delete c; return nullptr; }
  c->enc = reinterpret_cast<Encoder *>(c->handle);
  // one QP per picture (no delta-QP syntax); set on the object: the API's explicit-settings string goes
  // through std::stringstream, which is not safe inside a Python process that loaded another libstdc++
  c->enc->encoder_settings_.adaptive_qp = 0;
  c->enc->segment_header_->adaptive_qp = 0;
  c->width = width; c->height = height; c->bitdepth = bitdepth;
  return c;
}

void xref_conf_destroy(xref_conf *c) {
  if (!c) return;
  if (c->handle) c->api->encoder_destroy(c->handle);
  if (c->params) c->api->parameters_destroy(c->params);
  delete c;
}

// planes: tight 16-bit samples at the internal bit depth (input bit depth == internal bit depth)
int xref_conf_push_picture(xref_conf *c, const uint16_t *const planes[3]) {
  std::vector<uint8_t> bytes;
  for (int p = 0; p < 3; p++) {
    const size_t n = static_cast<size_t>(p ? c->width / 2 : c->width) * (p ? c->height / 2 : c->height);
    const uint8_t *b = reinterpret_cast<const uint8_t *>(planes[p]);
    bytes.insert(bytes.end(), b, b + 2 * n);
  }
  xvc_enc_nal_unit *units = nullptr;
  int n = 0;
  if (c->api->encoder_encode(c->handle, bytes.data(), &units, &n, nullptr) != XVC_ENC_OK) return -1;
  ConfCollect(c, units, n);
  return n;
}

int xref_conf_flush(xref_conf *c) {
  xvc_enc_nal_unit *units = nullptr;
  int n = 0, total = 0;
  while (c->api->encoder_flush(c->handle, &units, &n, nullptr) == XVC_ENC_OK && n > 0) {
    ConfCollect(c, units, n);
    total += n;
  }
  return total;
}

// What the outside encoder needs for picture `poc`: its original and the reconstruction of its
// (single) reference picture `ref_poc` at the internal bit depth, the picture QP and lambda, the
// segment's chroma QP mapping and the deblocking parameters.  info: qp, chroma table, offset u,
// offset v, deblock, beta offset, tc offset, picture type (0 bi, 1 uni).
int xref_conf_inter_inputs(xref_conf *c, int poc, int ref_poc, uint16_t *const orig[3], uint16_t *const ref_rec[3],
                           int32_t info[8], double *lambda) {
  auto pe = ConfFind(c, poc), re = ConfFind(c, ref_poc);
  if (!pe || !re) return -1;
  CopyOut(pe->orig_pic_.get(), orig);
  CopyOut(re->rec_pic_.get(), ref_rec);
  const Qp &qp = *pe->pic_data_->GetPicQp();
  const SegmentHeader &seg = *c->enc->segment_header_;
  info[0] = qp.GetQpRaw(YuvComponent::kY);
  info[1] = seg.chroma_qp_offset_table; info[2] = seg.chroma_qp_offset_u; info[3] = seg.chroma_qp_offset_v;
  info[4] = pe->pic_data_->GetDeblock() ? 1 : 0;
  info[5] = pe->pic_data_->GetBetaOffset(); info[6] = pe->pic_data_->GetTcOffset();
  info[7] = pe->pic_data_->GetPredictionType() == PicturePredictionType::kBi ? 0 : 1;
  *lambda = qp.GetLambda();
  return 0;
}

namespace {
struct ConfTreeBuilder {
  PictureData *pd;
  InterPrediction *inter;
  const xvcb200_cu *cus;
  int n_cus, next_cu = 0;
  const uint8_t *splits;
  int n_splits, next_split = 0;
  const int16_t *const *levels;
  int width;
  int pic_qp;
  bool ok = true;

  void Leaf(CodingUnit *cu) {
    if (next_cu >= n_cus) { ok = false; return; }
    const xvcb200_cu &d = cus[next_cu++];
    if (d.x != cu->GetPosX(YuvComponent::kY) || d.y != cu->GetPosY(YuvComponent::kY) ||
        d.w != cu->GetWidth(YuvComponent::kY) || d.h != cu->GetHeight(YuvComponent::kY) || d.ref_idx[0] != 0 ||
        d.ref_idx[1] >= 0) { ok = false; return; }
    cu->SetQp(pic_qp);
    cu->SetPredMode(PredictionMode::kInter);
    cu->SetSkipFlag(false);
    cu->SetMergeFlag(false);
    cu->SetInterDir(InterDir::kL0);
    cu->SetUseAffine(false);
    cu->SetUseLic(false);
    cu->SetFullpelMv(false);
    cu->SetRefIdx(0, RefPicList::kL0);
    cu->SetRefIdx(-1, RefPicList::kL1);
    const MotionVector mv(d.mv[0][0], d.mv[0][1]);
    cu->SetMv(mv, RefPicList::kL0);
    cu->SetMv(MotionVector(), RefPicList::kL1);
    // any predictor of the list codes the vector; take the cheaper one like EvalFinalMvpIdx would
    InterPredictorList mvp = inter->GetMvpList(*cu, RefPicList::kL0, 0);
    int best = 0;
    long best_cost = -1;
    for (int i = 0; i < static_cast<int>(mvp.size()); i++) {
      const MvDelta dd = mv - mvp[i];
      const long cost = std::labs(dd.x) + std::labs(dd.y);
      if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = i; }
    }
    cu->SetMvpIdx(best, RefPicList::kL0);
    cu->SetMvDelta(mv - mvp[best], RefPicList::kL0);
    cu->SetMvDelta(MvDelta(0, 0), RefPicList::kL1);
    bool any = false;
    for (int c = 0; c < 3; c++) {
      const YuvComponent comp = static_cast<YuvComponent>(c);
      const bool cbf = (d.flags & (XVCB200_CU_CBF_Y << c)) != 0;
      cu->SetCbf(comp, cbf);
      cu->SetTransformSkip(comp, false);
      cu->SetTransformType(comp, TransformType::kDefault, TransformType::kDefault);
      cu->SetDcCoeffOnly(comp, false);
      any |= cbf;
      const int w = cu->GetWidth(comp), h = cu->GetHeight(comp), x = cu->GetPosX(comp), y = cu->GetPosY(comp);
      const int pw = c ? width / 2 : width;
      CoeffBuffer cb = cu->GetCoeff(comp);
      for (int r = 0; r < h; r++)
        for (int q = 0; q < w; q++) cb.GetDataPtr()[r * cb.GetStride() + q] = cbf ? levels[c][static_cast<size_t>(y + r) * pw + x + q] : 0;
    }
    cu->SetRootCbf(any);
    cu->SetTransformFromSelectIdx(YuvComponent::kY, -1);
    pd->MarkUsedInPic(cu);
  }

  void Node(CodingUnit *cu) {
    if (!ok) return;
    if (next_split >= n_splits) { ok = false; return; }
    const int s = splits[next_split++];
    if (s == 0) { Leaf(cu); return; }
    // only splits the syntax can signal (CuWriter::WriteSplit)
    if (s == 1 && !(cu->GetDepth() < pd->GetMaxDepth(CuTree::Primary) && cu->GetBinaryDepth() == 0)) { ok = false; return; }
    if (s >= 2 && !cu->IsBinarySplitValid()) { ok = false; return; }
    cu->Split(static_cast<SplitType>(s));
    for (int i = 0; i < constants::kQuadSplit; i++)
      if (cu->GetSubCu(i)) Node(cu->GetSubCu(i));
  }
};
}  // namespace

// Re-codes picture `poc` (already coded once by the reference, which set up its headers and
// reference lists) from outside decisions: cus = leaf CUs in coding order with the chosen vector
// (list 0, reference index 0) and cbf flags, splits = the CU trees in pre-order (0 none, 1 quad,
// 2 horizontal, 3 vertical), levels = quantised coefficients (picture-shaped planes), rec = the
// reconstruction (after deblocking) the checksum is taken over.  Returns the NAL size or < 0.
int xref_conf_write_inter(xref_conf *c, int poc, const xvcb200_cu *cus, int n_cus, const uint8_t *splits, int n_splits,
                          const int16_t *const levels[3], const uint16_t *const rec[3]) {
  auto pe = ConfFind(c, poc);
  if (!pe) return -1;
  const bool trace = getenv("XREF_TRACE") != nullptr;
#define XT(msg) do { if (trace) { fprintf(stderr, "[conf] %s\n", msg); fflush(stderr); } } while (0)
  Encoder *enc = c->enc;
  XT("found picture");
  const SegmentHeader &segment = *enc->segment_header_;
  PictureData &pd = *pe->pic_data_;
  const Qp base_qp = *pd.GetPicQp();
  // reference lists first (PictureData::Init looks at them), as Encoder::EncodeOnePicture does
  ReferenceListSorter<PictureEncoder> sorter(segment, enc->prev_segment_header_->open_gop);
  sorter.Prepare(pe->GetPoc(), pd.GetTid(), pd.IsIntraPic(), enc->pic_encoders_, pd.GetRefPicLists(),
                 segment.leading_pictures);
  pd.Init(segment, base_qp, enc->encoder_settings_.adaptive_qp > 0);
  pd.SetUseLocalIlluminationCompensation(false);
  XT("init done");

  XT("ref lists prepared");
  BitWriter &bw = pe->bit_writer_;
  bw.Clear();
  pe->WriteHeader(segment, pd, static_cast<PicNum>(segment.max_sub_gop_length), pe->GetBufferFlag(), &bw);
  SyntaxWriter writer(base_qp, pd.GetPredictionType(), &bw);
  IntraPrediction intra_pred(c->bitdepth);
  CuWriter cu_writer(pd, &intra_pred);
  InterPrediction inter(Simd(1, c->bitdepth).inter_prediction, *pe->rec_pic_, c->bitdepth);
  ConfTreeBuilder tb;
  tb.pd = &pd; tb.inter = &inter; tb.cus = cus; tb.n_cus = n_cus; tb.splits = splits; tb.n_splits = n_splits;
  tb.levels = levels; tb.width = c->width; tb.pic_qp = base_qp.GetQpRaw(YuvComponent::kY);
  const int num_ctus = pd.GetNumberOfCtu();
  for (int rsaddr = 0; rsaddr < num_ctus; rsaddr++) {
    CodingUnit *ctu = pd.GetCtu(CuTree::Primary, rsaddr);
    ctu->SetQp(tb.pic_qp);
    XT("ctu: build");
    tb.Node(ctu);
    if (!tb.ok) return -2;
    XT("ctu: write");
    cu_writer.WriteCtu(ctu, &pd, &writer);
    if (Restrictions::Get().disable_ext_implicit_last_ctu) writer.WriteEndOfSlice(false);
  }
  if (tb.next_cu != n_cus || tb.next_split != n_splits) return -3;
  XT("ctus written");
  writer.Finish();
  CopyIn(pe->rec_pic_.get(), rec);
  pe->rec_pic_->PadBorder();
  pd.GetRefPicLists()->ZeroOutReferences();
  pe->WriteChecksum(segment, &bw, segment.checksum_mode);
  const std::vector<uint8_t> *bytes = bw.GetBytes();
  for (auto &nal : c->nals)
    if (nal.type != 16 && static_cast<int>(nal.poc) == poc) { nal.bytes = *bytes; return static_cast<int>(bytes->size()); }
  return -4;
}

// The bitstream in the file framing of the reference applications (4-byte little-endian NAL
// size before every NAL, encoder_app.cc:494-516).
size_t xref_conf_bitstream(xref_conf *c, uint8_t *out, size_t cap) {
  size_t need = 0;
  for (auto &nal : c->nals) need += 4 + nal.bytes.size();
  if (!out || cap < need) return need;
  size_t off = 0;
  for (auto &nal : c->nals) {
    const uint32_t sz = static_cast<uint32_t>(nal.bytes.size());
    for (int i = 0; i < 4; i++) out[off++] = static_cast<uint8_t>((sz >> (8 * i)) & 0xff);
    std::memcpy(out + off, nal.bytes.data(), sz);
    off += sz;
  }
  return need;
}

int xref_num_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? static_cast<int>(n) : 1;
}

}  // extern "C"
