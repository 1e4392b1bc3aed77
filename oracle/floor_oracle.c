/*
 * oracle/floor_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, float64) of the reference's Phase-3 / counting arithmetic.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Pinned against the reference itself: tests/golden/make_golden.py imports the reference's Python
 * classes from /root/reference and stores their outputs; tests/test_oracle.py checks this file
 * against those fixtures (and against the reference's own known-answer tests).
 *
 * Every function cites the reference lines it follows.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* src/transform/homography.py:166-169 — foot point of an (x, y, w, h) box */
static void foot_point(const double* b, double* fx, double* fy) {
  *fx = b[0] + b[2] / 2;
  *fy = b[1] + b[3];
}

/* src/transform/homography.py:172-175 — [X,Y,W] = H·[x,y,1]; (X/W, Y/W).
 * NumPy evaluates the 3-term dot products through BLAS; the summation order is k = 0,1,2. */
static void project(const double* H, double x, double y, double* px, double* py) {
  volatile double X = H[0] * x; X += H[1] * y; X += H[2] * 1.0;
  volatile double Y = H[3] * x; Y += H[4] * y; Y += H[5] * 1.0;
  volatile double W = H[6] * x; W += H[7] * y; W += H[8] * 1.0;
  *px = X / W;
  *py = Y / W;
}

/* src/transform/homography.py:150-197 transform_batch (is_bbox=1) / :105-133 transform_pixel (is_bbox=0) */
void oracle_transform(const double* H, const double* rows, int64_t n, int is_bbox, double width_px,
                      double height_px, double sx_mm, double sy_mm, double* floor_px, double* floor_mm,
                      uint8_t* within) {
  for (int64_t i = 0; i < n; ++i) {
    double x, y;
    if (is_bbox) foot_point(rows + 4 * i, &x, &y);
    else { x = rows[2 * i]; y = rows[2 * i + 1]; }
    double px, py;
    project(H, x, y, &px, &py);
    if (floor_px) { floor_px[2 * i] = px; floor_px[2 * i + 1] = py; }
    if (floor_mm) { floor_mm[2 * i] = px * sx_mm; floor_mm[2 * i + 1] = py * sy_mm; }       /* :183-186 */
    if (within) within[i] = (0 <= px && px < width_px && 0 <= py && py < height_px) ? 1 : 0;  /* :181 */
  }
}

/* src/zone/zone_classifier.py:162-197 _point_in_polygon — ray casting with the half-open rule */
int oracle_point_in_polygon(double x, double y, const double* poly, int n) {
  int inside = 0;
  double p1x = poly[0], p1y = poly[1];
  volatile double xinters = 0.0;
  for (int i = 1; i <= n; ++i) {
    const double p2x = poly[2 * (i % n)], p2y = poly[2 * (i % n) + 1];
    const double ymin = p1y < p2y ? p1y : p2y, ymax = p1y < p2y ? p2y : p1y;
    const double xmax = p1x < p2x ? p2x : p1x;
    if (y > ymin && y <= ymax && x <= xmax) {
      if (p1y != p2y) {
        volatile double a = y - p1y, b = p2x - p1x;
        volatile double c = a * b;
        volatile double d = p2y - p1y;
        volatile double e = c / d;
        xinters = e + p1x;
      }
      if (p1x == p2x || x <= xinters) inside = !inside;
    }
    p1x = p2x; p1y = p2y;
  }
  return inside;
}

/* src/zone/zone_classifier.py:114-149 classify.
 * zone_mask[i]: bit z set <=> point i inside zone z (declaration order) — the allow_overlap=True answer.
 * zone_idx[i] : argmin (priority or +inf, declaration order) over containing zones, -1 if none — the
 *               allow_overlap=False answer (:138-146). */
void oracle_classify(const double* pts, int64_t n, const double* verts, const int32_t* offs, const double* prio,
                     int Z, int32_t* zone_idx, uint64_t* zone_mask) {
  for (int64_t i = 0; i < n; ++i) {
    uint64_t m = 0;
    int best = -1;
    double bp = INFINITY;
    for (int z = 0; z < Z; ++z) {
      if (oracle_point_in_polygon(pts[2 * i], pts[2 * i + 1], verts + 2 * offs[z], offs[z + 1] - offs[z])) {
        if (z < 64) m |= 1ull << z;
        const double p = isnan(prio[z]) ? INFINITY : prio[z];
        if (best < 0 || p < bp) { best = z; bp = p; }  /* ties keep the earlier declaration */
      }
    }
    if (zone_idx) zone_idx[i] = best;
    if (zone_mask) zone_mask[i] = m;
  }
}

/* src/aggregation/aggregator.py:52-75 get_zone_counts, dense form: hist[T][Z+1], column Z = "unclassified".
 * Exactly one of zone_idx / zone_mask is given. */
void oracle_count(const int32_t* zone_idx, const uint64_t* zone_mask, const int32_t* slot, int64_t n, int Z, int T,
                  int64_t* hist) {
  for (int64_t i = 0; i < n; ++i) {
    const int s = slot ? slot[i] : 0;
    if (s < 0 || s >= T) continue;
    int64_t* row = hist + (int64_t)s * (Z + 1);
    if (zone_idx) {
      row[(zone_idx[i] >= 0 && zone_idx[i] < Z) ? zone_idx[i] : Z] += 1;
    } else {
      uint64_t m = zone_mask[i];
      if (m == 0) row[Z] += 1;
      for (int z = 0; z < Z && z < 64; ++z) if ((m >> z) & 1ull) row[z] += 1;
    }
  }
}

/* Distance from each point to the nearest polygon edge (for the north star's "not within 1e-4 px of an
 * edge" exemption; no reference counterpart). */
void oracle_min_edge_distance(const double* pts, int64_t n, const double* verts, const int32_t* offs, int Z,
                              double* dist) {
  for (int64_t i = 0; i < n; ++i) {
    const double x = pts[2 * i], y = pts[2 * i + 1];
    double best = INFINITY;
    for (int z = 0; z < Z; ++z) {
      const int o = offs[z], nv = offs[z + 1] - o;
      for (int k = 0; k < nv; ++k) {
        const double ax = verts[2 * (o + k)], ay = verts[2 * (o + k) + 1];
        const double bx = verts[2 * (o + (k + 1) % nv)], by = verts[2 * (o + (k + 1) % nv) + 1];
        const double dx = bx - ax, dy = by - ay, l2 = dx * dx + dy * dy;
        double t = l2 > 0 ? ((x - ax) * dx + (y - ay) * dy) / l2 : 0.0;
        t = t < 0 ? 0 : (t > 1 ? 1 : t);
        const double ex = ax + t * dx - x, ey = ay + t * dy - y;
        const double d = sqrt(ex * ex + ey * ey);
        if (d < best) best = d;
      }
    }
    dist[i] = best;
  }
}

/* The whole Phase-3 + count path on points, as bench.py's CPU baseline times it (single thread). */
void oracle_project_classify_count(const double* H, const float* pts_f32, int64_t n, const double* verts,
                                   const int32_t* offs, const double* prio, int Z, int32_t* zone_idx,
                                   int64_t* hist) {
  for (int64_t i = 0; i < n; ++i) {
    double p[2];
    project(H, (double)pts_f32[2 * i], (double)pts_f32[2 * i + 1], &p[0], &p[1]);
    int32_t zi;
    oracle_classify(p, 1, verts, offs, prio, Z, &zi, 0);
    if (zone_idx) zone_idx[i] = zi;
    if (hist) hist[zi >= 0 ? zi : Z] += 1;
  }
}
