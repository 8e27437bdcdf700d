// RPN decode + rotated-box non-maximum suppression on sm_100a: what rpnToRegion does to the two
// tensors model.predict returns (reference rpnToRegion.py:18-164, IoU from serialize_data.py:140-181):
//
//     A[:, a, b, i]  = anchor i at the centre of output cell (a, b)            rpnToRegion.py:137-144
//     box            = applyRegrssion(anchor, t)                               :77-88   (x,y,z linear, l,w,h by exp, yaw added)
//     probInfo/boxInfo = anchor-major flattening  i*outX*outY + a*outY + b     :150-152
//     nonMaxSuppressionFast(boxInfo, probInfo, maxBoxes=20, overlapThresh=0.)  :18-74, :163
//
// decode_kernel: one thread per candidate, HBM-bound (16 floats read, 7 doubles + 1 float written per position/anchor).
// Arithmetic as numpy does it: float32 regressions widened to float64, t*size rounded then + anchor rounded (no FMA);
// exp in float32 (np.exp of a float32 array), widened, times the float64 anchor size.
//
// nms_kernel: greedy suppression is a chain of dependent rounds (pick the best survivor, delete everything it overlaps),
// at most maxBoxes + 1 of them, each a max-reduction plus one overlap test per survivor. One thread-block CLUSTER of 8
// CTAs takes one sample: a CTA keeps its eighth of the candidates (score, centre, bounding radius as float32, survivor
// bits) in shared memory for the whole run; per round the CTAs publish their local best, meet at ONE cluster barrier,
// read each other's slot through distributed shared memory and test their own survivors against the pick. The float64
// box of a candidate is only fetched when the float32 bounding circles (with a safety margin) touch; a float32
// separating-axis screen then settles the pairs that are apart, or (for overlapThresh = 0, the reference's call site)
// overlapping, by more than its own error bound; what remains gets the overlap the reference computes: polygon corners as boxToShapely builds them (serialize_data.py:151-163), area of the intersection
// polygon (Sutherland-Hodgman, standing where shapely's .intersection().area stood), z overlap with the reference's
// full-height extents (z -/+ h, :146-147), iou = intersect / (vol1 + vol2 - intersect) > overlapThresh. Every float64
// operation is an explicit round-to-nearest intrinsic so that no FMA contraction separates it from the CPU restatement.
// Deviations a maintainer must know: ties in score go to the larger flat index (np.argsort's quicksort leaves tie
// order unspecified); NaN scores are picked last (numpy sorts them to the end, i.e. picks them first).
#include <cooperative_groups.h>

#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace lisec {

namespace {

constexpr int kNmsCluster = 8;
constexpr int kNmsThreads = 512;
constexpr int kNmsMaxSlice = 12288;  // candidates per CTA: 16 B each + 1 bit -> 198 KB of shared memory

struct DecodeParams {
  int out_x, out_y, n_anchors;
  double cell_x, cell_y, anchor_z;
  double anchors[LISEC_MAX_ANCHORS][4];  // l, w, h, yaw  (Constants.py:17)
  long long prob_pitch, reg_pitch, prob_batch, reg_batch;  // in floats
};

__global__ void __launch_bounds__(256)
    decode_kernel(const float* __restrict__ prob, const float* __restrict__ reg, const __grid_constant__ DecodeParams P,
                  int batch, double* __restrict__ boxes, float* __restrict__ scores) {
  pdl_launch_dependents();
  const int per = P.out_x * P.out_y;
  const long long n = (long long)per * P.n_anchors;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (gid >= n * batch) return;
  const int s = (int)(gid / n);
  const int c = (int)(gid - (long long)s * n);  // flat candidate index: anchor-major (rpnToRegion.py:150-152)
  const int i = c / per, pos = c - i * per, a = pos / P.out_y, b = pos - a * P.out_y;
  const float* t = reg + (size_t)s * P.reg_batch + (size_t)pos * P.reg_pitch + 7 * i;
  const double ax = __dadd_rn(__dmul_rn((double)a, P.cell_x), P.cell_x / 2);  // X.T * voxelXSize + voxelXSize / 2 (:137)
  const double ay = __dadd_rn(__dmul_rn((double)b, P.cell_y), P.cell_y / 2);
  const double l = P.anchors[i][0], w = P.anchors[i][1], h = P.anchors[i][2];
  double* o = boxes + ((size_t)s * n + c) * 7;
  o[0] = __dadd_rn(__dmul_rn((double)__ldcg(t + 0), l), ax);  // tx * l + x (:80)
  o[1] = __dadd_rn(__dmul_rn((double)__ldcg(t + 1), w), ay);
  o[2] = __dadd_rn(__dmul_rn((double)__ldcg(t + 2), h), P.anchor_z);
  o[3] = __dmul_rn((double)expf(__ldcg(t + 3)), l);  // np.exp(tl) * l: the exp is float32's (:83)
  o[4] = __dmul_rn((double)expf(__ldcg(t + 4)), w);
  o[5] = __dmul_rn((double)expf(__ldcg(t + 5)), h);
  o[6] = __dadd_rn((double)__ldcg(t + 6), P.anchors[i][3]);
  scores[(size_t)s * n + c] = __ldcg(prob + (size_t)s * P.prob_batch + (size_t)pos * P.prob_pitch + i);
}

// ---- geometry (serialize_data.py:140-181) ---------------------------------------------------------------------------
struct Quad {
  double x[4], y[4];
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// boxToShapely (:151-163): [topRight, botRight, botLeft, topLeft]
__device__ void box_corners(const double* bx, Quad& q) {
  const double theta = bx[6], hl = __ddiv_rn(bx[3], 2.0), hw = __ddiv_rn(bx[4], 2.0);
  const double c = cos(theta), s = sin(theta);
  const double rrx = dadd(bx[0], dmul(c, hw)), rry = dsub(bx[1], dmul(s, hw));
  const double rlx = dsub(bx[0], dmul(c, hw)), rly = dadd(bx[1], dmul(s, hw));
  q.x[0] = dadd(rrx, dmul(s, hl)); q.y[0] = dadd(rry, dmul(c, hl));
  q.x[1] = dsub(rrx, dmul(s, hl)); q.y[1] = dsub(rry, dmul(c, hl));
  q.x[2] = dsub(rlx, dmul(s, hl)); q.y[2] = dsub(rly, dmul(c, hl));
  q.x[3] = dadd(rlx, dmul(s, hl)); q.y[3] = dadd(rly, dmul(c, hl));
}

__device__ __forceinline__ double cross2(double ax, double ay, double bx, double by, double cx, double cy) {
  return dsub(dmul(dsub(bx, ax), dsub(cy, ay)), dmul(dsub(by, ay), dsub(cx, ax)));  // (B - A) x (C - A)
}

// Area of subject ∩ clip, both convex quadrilaterals (oracle/decode_oracle.py: quad_intersection_area, same order of
// operations).
__device__ double quad_intersection_area(const Quad& subj, const Quad& clip) {
  double sa = 0.0;
  for (int i = 0; i < 4; ++i) {
    const int j = (i + 1) & 3;
    sa = dadd(sa, dsub(dmul(clip.x[i], clip.y[j]), dmul(clip.x[j], clip.y[i])));
  }
  if (sa == 0.0) return 0.0;
  const double orient = sa > 0.0 ? 1.0 : -1.0;
  double px[12], py[12], qx[12], qy[12];
  int n = 4;
  for (int i = 0; i < 4; ++i) { px[i] = subj.x[i]; py[i] = subj.y[i]; }
  for (int e = 0; e < 4; ++e) {
    const double ax = clip.x[e], ay = clip.y[e], bx = clip.x[(e + 1) & 3], by = clip.y[(e + 1) & 3];
    int m = 0;
    for (int k = 0; k < n; ++k) {
      const int k1 = k + 1 == n ? 0 : k + 1;
      const double sc = dmul(orient, cross2(ax, ay, bx, by, px[k], py[k]));
      const double sd = dmul(orient, cross2(ax, ay, bx, by, px[k1], py[k1]));
      const bool in_c = sc >= 0.0, in_d = sd >= 0.0;
      if (in_c) { qx[m] = px[k]; qy[m] = py[k]; ++m; }
      if (in_c != in_d) {
        const double t = __ddiv_rn(sc, dsub(sc, sd));
        qx[m] = dadd(px[k], dmul(t, dsub(px[k1], px[k])));
        qy[m] = dadd(py[k], dmul(t, dsub(py[k1], py[k])));
        ++m;
      }
    }
    n = m;
    if (n == 0) return 0.0;
    for (int k = 0; k < n; ++k) { px[k] = qx[k]; py[k] = qy[k]; }
  }
  double a2 = 0.0;
  for (int k = 0; k < n; ++k) {
    const int k1 = k + 1 == n ? 0 : k + 1;
    a2 = dadd(a2, dsub(dmul(px[k], py[k1]), dmul(px[k1], py[k])));
  }
  return dmul(0.5, fabs(a2));
}

struct NmsParams {
  int n;             // candidates per sample
  int max_boxes;     // the loop stops once len(pick) > max_boxes (rpnToRegion.py:71)
  double thresh;     // overlapThresh
  double margin_x, margin_y, limit_x, limit_y;  // range test of :55-58: x - mx < 0 or x + mx > limit_x or ...
  int slice;         // candidates per CTA
};

struct Best {
  float score;
  int idx;
};

__device__ __forceinline__ bool better(float s, int i, float bs, int bi) { return s > bs || (s == bs && i > bi); }

__global__ void __cluster_dims__(kNmsCluster, 1, 1) __launch_bounds__(kNmsThreads, 1)
    nms_kernel(const double* __restrict__ boxes, const float* __restrict__ scores, const __grid_constant__ NmsParams P,
               int* __restrict__ picks, int* __restrict__ n_picks, double* __restrict__ out_boxes,
               float* __restrict__ out_scores) {
  pdl_launch_dependents();
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char smem[];
  float* s_score = reinterpret_cast<float*>(smem);
  float* s_cx = s_score + P.slice;
  float* s_cy = s_cx + P.slice;
  float* s_r = s_cy + P.slice;
  unsigned* s_alive = reinterpret_cast<unsigned*>(s_r + P.slice);  // (slice + 31) / 32 words
  __shared__ Best s_slot[2];       // this CTA's best survivor, double-buffered by round parity
  __shared__ Best s_warp[kNmsThreads / 32];
  __shared__ double s_pick[7];
  __shared__ Quad s_quad;

  const int rank = (int)cluster.block_rank();
  const int sample = blockIdx.x / kNmsCluster;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = rank * P.slice, cnt = max(0, min(P.slice, P.n - c0));
  const double* bx = boxes + (size_t)sample * P.n * 7;
  const float* sc = scores + (size_t)sample * P.n;
  const int words = (P.slice + 31) >> 5;
  pdl_wait();

  for (int w = tid; w < words; w += kNmsThreads) {
    const int left = cnt - 32 * w;
    s_alive[w] = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
  }
  for (int k = tid; k < cnt; k += kNmsThreads) {
    const double* b = bx + (size_t)(c0 + k) * 7;
    const float s = sc[c0 + k];
    s_score[k] = s == s ? s : -INFINITY;  // NaN scores go last
    s_cx[k] = (float)b[0];
    s_cy[k] = (float)b[1];
    // bounding circle of the footprint, rounded up
    s_r[k] = (float)(0.5 * sqrt(b[3] * b[3] + b[4] * b[4])) * 1.000001f + 1e-6f;
  }
  __syncthreads();

  int n_pick = 0;
  for (int round = 0;; ++round) {
    // ---- local best survivor ----
    Best best = {-INFINITY, -1};
    for (int k = tid; k < cnt; k += kNmsThreads)
      if ((s_alive[k >> 5] >> (k & 31)) & 1u) {
        const float s = s_score[k];
        if (best.idx < 0 || better(s, c0 + k, best.score, best.idx)) best = {s, c0 + k};
      }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, best.score, d);
      const int oi = __shfl_xor_sync(0xffffffffu, best.idx, d);
      if (oi >= 0 && (best.idx < 0 || better(os, oi, best.score, best.idx))) best = {os, oi};
    }
    if (lane == 0) s_warp[warp] = best;
    __syncthreads();
    if (warp == 0) {
      best = lane < kNmsThreads / 32 ? s_warp[lane] : Best{-INFINITY, -1};
#pragma unroll
      for (int d = 8; d > 0; d >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, best.score, d);
        const int oi = __shfl_xor_sync(0xffffffffu, best.idx, d);
        if (oi >= 0 && (best.idx < 0 || better(os, oi, best.score, best.idx))) best = {os, oi};
      }
      if (lane == 0) s_slot[round & 1] = best;
    }
    cluster.sync();  // every CTA's slot of this round is published (and last round's reads are long done)
    Best g = {-INFINITY, -1};
    for (int r = 0; r < kNmsCluster; ++r) {
      const Best o = *cluster.map_shared_rank(&s_slot[round & 1], r);
      if (o.idx >= 0 && (g.idx < 0 || better(o.score, o.idx, g.score, g.idx))) g = o;
    }
    if (g.idx < 0) break;  // nothing left: the same decision in all CTAs of the cluster
    // ---- the pick ----
    if (tid < 7) s_pick[tid] = bx[(size_t)g.idx * 7 + tid];
    if (tid == 32 && g.idx >= c0 && g.idx < c0 + cnt) atomicAnd(&s_alive[(g.idx - c0) >> 5], ~(1u << ((g.idx - c0) & 31)));
    __syncthreads();
    if (tid == 0) box_corners(s_pick, s_quad);
    if (rank == 0 && tid < 7) out_boxes[((size_t)sample * (P.max_boxes + 1) + n_pick) * 7 + tid] = s_pick[tid];
    if (rank == 0 && tid == 7) {
      picks[(size_t)sample * (P.max_boxes + 1) + n_pick] = g.idx;
      out_scores[(size_t)sample * (P.max_boxes + 1) + n_pick] = sc[g.idx];
    }
    ++n_pick;
    __syncthreads();
    // ---- delete what the pick overlaps, and what is out of range (rpnToRegion.py:53-66) ----
    const float pcx = (float)s_pick[0], pcy = (float)s_pick[1];
    const float pr = (float)(0.5 * sqrt(s_pick[3] * s_pick[3] + s_pick[4] * s_pick[4])) * 1.000001f + 1e-6f;
    const double vol_p = dmul(dmul(s_pick[3], s_pick[4]), s_pick[5]);
    const float plf = (float)s_pick[3], pwf = (float)s_pick[4];
    float psin, pcos;
    sincosf((float)s_pick[6], &psin, &pcos);
    const bool pick_positive = s_pick[3] > 0.0 && s_pick[4] > 0.0 && s_pick[5] > 0.0;
    for (int k = tid; k < cnt; k += kNmsThreads) {
      if (!((s_alive[k >> 5] >> (k & 31)) & 1u)) continue;
      const double* b = bx + (size_t)(c0 + k) * 7;
      bool del = false;
      const float fx = s_cx[k], fy = s_cy[k];
      // the float32 copies decide only when they are far from the limits; otherwise the float64 values do
      const float mx = (float)P.margin_x, my = (float)P.margin_y, lx = (float)P.limit_x, ly = (float)P.limit_y;
      const float slack = 1e-3f;
      if (fx - mx < -slack || fx + mx > lx + slack || fy - my < -slack || fy + my > ly + slack) {
        del = true;
      } else if (fx - mx < slack || fx + mx > lx - slack || fy - my < slack || fy + my > ly - slack) {
        const double x = b[0], y = b[1];
        del = dsub(x, P.margin_x) < 0.0 || dadd(x, P.margin_x) > P.limit_x || dsub(y, P.margin_y) < 0.0 ||
              dadd(y, P.margin_y) > P.limit_y;
      }
      if (!del) {
        const float dx = fx - pcx, dy = fy - pcy, rr = s_r[k] + pr;
        if (dx * dx + dy * dy <= rr * rr * 1.0001f + 1e-4f) {  // circles touch
          // float32 separating-axis screen (boxToShapely's rectangle: half-width along (cos, -sin), half-length along
          // (sin, cos)): a gap or an overlap deeper than the float32 error bound decides; the rest goes the exact way.
          const float cl = (float)b[3], cw = (float)b[4], cyaw = (float)b[6];
          float cs, cc;
          sincosf(cyaw, &cs, &cc);
          const float ax[4][2] = {{pcos, -psin}, {psin, pcos}, {cc, -cs}, {cs, cc}};
          float worst = -INFINITY;
          bool decided_apart = false, finite = true;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const float dist = fabsf(dx * ax[a][0] + dy * ax[a][1]);
            const float rp = 0.5f * (pwf * fabsf(ax[0][0] * ax[a][0] + ax[0][1] * ax[a][1]) +
                                     plf * fabsf(ax[1][0] * ax[a][0] + ax[1][1] * ax[a][1]));
            const float rc = 0.5f * (cw * fabsf(ax[2][0] * ax[a][0] + ax[2][1] * ax[a][1]) +
                                     cl * fabsf(ax[3][0] * ax[a][0] + ax[3][1] * ax[a][1]));
            const float gap = dist - (rp + rc), t = 1e-3f + 8e-6f * (dist + rp + rc + fabsf(fx) + fabsf(fy));
            finite = finite && gap == gap && t == t;
            decided_apart = decided_apart || gap > t;
            worst = fmaxf(worst, gap + t);  // < 0 on every axis: overlap deeper than the error bound
          }
          const bool deep = finite && worst < 0.f && !decided_apart;  // NaN anywhere: exact path
          const double ch = b[5], cz = b[2];
          if (decided_apart && finite) {
            del = false;  // area 0 -> iou = 0, never above a threshold >= 0
          } else if (deep && P.thresh == 0.0 && pick_positive && cl > 0.f && cw > 0.f && ch > 0.0) {
            // area > 0 and all sizes > 0: intersect has the sign of the z overlap, union >= 0 (0 only for coincident
            // boxes, where the reference's division yields +inf): iou > 0 exactly when the z extents overlap
            const double bot = fmax(dsub(s_pick[2], s_pick[5]), dsub(cz, ch));
            const double top = fmin(dadd(s_pick[2], s_pick[5]), dadd(cz, ch));
            del = dsub(top, bot) > 0.0;
          } else {
            double cand[7];
#pragma unroll
            for (int q = 0; q < 7; ++q) cand[q] = b[q];
            Quad cq;
            box_corners(cand, cq);
            // calculateIntersection(lastBox, box): box1 = the pick (rpnToRegion.py:64, serialize_data.py:140-148)
            const double area = quad_intersection_area(s_quad, cq);
            const double bot = fmax(dsub(s_pick[2], s_pick[5]), dsub(cand[2], cand[5]));
            const double top = fmin(dadd(s_pick[2], s_pick[5]), dadd(cand[2], cand[5]));
            const double inter = dmul(dsub(top, bot), area);
            const double uni = dsub(dadd(vol_p, dmul(dmul(cand[3], cand[4]), cand[5])), inter);
            const double iou = __ddiv_rn(inter, uni);
            del = iou > P.thresh;
          }
        }
      }
      if (del) atomicAnd(&s_alive[k >> 5], ~(1u << (k & 31)));
    }
    __syncthreads();
    if (n_pick > P.max_boxes) break;  // `if len(pick) > maxBoxes: break` (:71)
  }
  if (rank == 0 && tid == 0) n_picks[sample] = n_pick;
  for (int k = n_pick + tid; rank == 0 && k <= P.max_boxes; k += kNmsThreads)
    picks[(size_t)sample * (P.max_boxes + 1) + k] = -1;
  cluster.sync();  // no CTA may exit while a peer can still read its slot
}

thread_local char g_decode_error[256] = "";

int32_t decode_fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_decode_error, sizeof(g_decode_error), fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace

}  // namespace lisec

using namespace lisec;

extern "C" {

const char* lisec_decode_last_error(void) { return g_decode_error; }

int32_t lisec_rpn_decode(const lisec_rpn_desc* d, const float* prob, int64_t prob_pitch, int64_t prob_batch_stride,
                         const float* regress, int64_t reg_pitch, int64_t reg_batch_stride, int32_t batch,
                         double* boxes, float* scores, void* stream) {
  if (!d || !prob || !regress || !boxes || !scores) return decode_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (d->out_x <= 0 || d->out_y <= 0 || d->n_anchors <= 0 || d->n_anchors > LISEC_MAX_ANCHORS || batch < 0)
    return decode_fail(LISEC_ERR_BAD_CONFIG, "out_x, out_y > 0 and 1 <= n_anchors <= %d", LISEC_MAX_ANCHORS);
  if (prob_pitch < d->n_anchors || reg_pitch < 7 * d->n_anchors)
    return decode_fail(LISEC_ERR_BAD_ARG, "pitches smaller than the channels they hold");
  if (batch == 0) return LISEC_OK;
  DecodeParams p;
  p.out_x = d->out_x; p.out_y = d->out_y; p.n_anchors = d->n_anchors;
  p.cell_x = d->cell_x; p.cell_y = d->cell_y; p.anchor_z = d->anchor_z;
  for (int i = 0; i < d->n_anchors; ++i)
    for (int k = 0; k < 4; ++k) p.anchors[i][k] = d->anchors[i][k];
  p.prob_pitch = prob_pitch; p.reg_pitch = reg_pitch; p.prob_batch = prob_batch_stride; p.reg_batch = reg_batch_stride;
  const long long total = (long long)d->out_x * d->out_y * d->n_anchors * batch;
  cudaError_t e = launch_pdl(decode_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                             static_cast<cudaStream_t>(stream), prob, regress, p, (int)batch, boxes, scores);
  if (e != cudaSuccess) return decode_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_nms_rotated(const lisec_nms_desc* d, const double* boxes, const float* scores, int32_t n, int32_t batch,
                          int32_t* picks, int32_t* n_picks, double* out_boxes, float* out_scores, void* stream) {
  if (!d || !picks || !n_picks || !out_boxes || !out_scores) return decode_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (n < 0 || batch < 0 || d->max_boxes < 0 || d->max_boxes > 4095)
    return decode_fail(LISEC_ERR_BAD_ARG, "n, batch >= 0 and 0 <= max_boxes <= 4095");
  if (!(d->overlap_thresh >= 0.0))
    return decode_fail(LISEC_ERR_UNSUPPORTED, "overlap_thresh must be >= 0 (disjoint boxes are never tested)");
  if (n > kNmsCluster * kNmsMaxSlice)
    return decode_fail(LISEC_ERR_CAPACITY, "at most %d candidates per sample", kNmsCluster * kNmsMaxSlice);
  if (batch == 0) return LISEC_OK;
  if (n > 0 && (!boxes || !scores)) return decode_fail(LISEC_ERR_BAD_ARG, "null argument");
  NmsParams p;
  p.n = n; p.max_boxes = d->max_boxes; p.thresh = d->overlap_thresh;
  p.margin_x = d->margin_x; p.margin_y = d->margin_y; p.limit_x = d->limit_x; p.limit_y = d->limit_y;
  p.slice = (n + kNmsCluster - 1) / kNmsCluster;
  p.slice = (p.slice + 31) & ~31;
  if (p.slice == 0) p.slice = 32;
  const size_t smem = (size_t)p.slice * 16 + (size_t)(p.slice / 32) * 4;
  cudaError_t e = cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return decode_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  // __cluster_dims__ fixes the cluster shape; the grid is batch clusters
  nms_kernel<<<dim3((unsigned)(batch * kNmsCluster)), dim3(kNmsThreads), smem, static_cast<cudaStream_t>(stream)>>>(
      boxes, scores, p, picks, n_picks, out_boxes, out_scores);
  e = cudaGetLastError();
  if (e != cudaSuccess) return decode_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

}  // extern "C"
