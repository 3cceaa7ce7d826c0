// geometric_coeffs.cuh -- analytic direct-beam coefficients of a 3-stream box with sloped top and bottom faces
// (-pprts_geometric_coeffs; src/geometric_coeffs.F90:34-404 with m_intersection's hit_plane / line_intersection_3d,
// src/intersection.fypp:79-91, :395-434, and the Heron areas of src/helper_functions.fypp:2116-2136, :3087-3092).
// Every source face is projected along the sun onto the faces opposite to it; the lit areas, weighted with Beer-Lambert
// along the beam (closed form on the rectangular part, num_intervals-point midpoint rule on the triangular part),
// are the dir2dir coefficients.  fp64 throughout; one thread computes the 9 coefficients of its cell.
// Flat index dst*3 + src as everywhere in this library.  Vertex order a,b,c,d (face at the lower z), e,f,g,h.
#pragma once
#include <cfloat>
#include <cmath>

typedef double gv3[3];
__device__ static void g_sub(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
__device__ static double g_dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ static double g_norm(const double *a) { return sqrt(g_dot(a, a)); }
__device__ static void g_axpy(const double *o, double c, const double *d, double *out) { out[0] = o[0] + c * d[0]; out[1] = o[1] + c * d[1]; out[2] = o[2] + c * d[2]; }
__device__ static void g_copy(const double *a, double *o) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
/* compute_normal_3d, helper_functions.fypp:1960-1974 */
__device__ static void g_normal(const double *p1, const double *p2, const double *p3, double *n) {
  gv3 U, V; double l;
  g_sub(p2, p1, U); g_sub(p3, p1, V);
  n[0] = U[1] * V[2] - U[2] * V[1]; n[1] = U[2] * V[0] - U[0] * V[2]; n[2] = U[0] * V[1] - U[1] * V[0];
  l = g_norm(n); n[0] /= l; n[1] /= l; n[2] /= l;
}
/* hit_plane, intersection.fypp:79-91 */
__device__ static double g_hit_plane(const double *p_loc, const double *p_dir, const double *po, const double *pn) {
  double discr = g_dot(p_dir, pn); gv3 d;
  if (discr <= DBL_EPSILON && discr > -DBL_EPSILON) return sqrt(DBL_MAX);
  g_sub(po, p_loc, d);
  return g_dot(d, pn) / discr;
}
/* triangle_area_by_vertices / quadrangle_area_by_vertices, helper_functions.fypp:2116-2136, :3087-3092 (Heron) */
__device__ static double g_tri(const double *v1, const double *v2, const double *v3) {
  gv3 d; double e1, e2, e3, p2;
  g_sub(v2, v1, d); e1 = g_norm(d); g_sub(v3, v2, d); e2 = g_norm(d); g_sub(v1, v3, d); e3 = g_norm(d);
  p2 = (e1 + e2 + e3) / 2;
  return sqrt(fmax(p2 * (p2 - e1) * (p2 - e2) * (p2 - e3), 0.0));
}
__device__ static double g_quad(const double *v1, const double *v2, const double *v3, const double *v4) { return g_tri(v1, v2, v3) + g_tri(v1, v3, v4); }
/* line_intersection_3d, intersection.fypp:395-434 */
__device__ static int g_line_intersection(const double *p1, const double *dir1, const double *p3, const double *dir2, double *c1, double *c2) {
  gv3 a, b, p13; double na, nb, nd, den, d2121, d4343, d4321, d1343, d1321;
  const double limit = sqrt(DBL_EPSILON);
  g_copy(dir1, a); g_copy(dir2, b); /* p2 - p1 = direction1, p4 - p3 = direction2 */
  na = g_norm(a); nb = g_norm(b);
  nd = g_dot(a, b) / (na * nb);
  den = 1.0 * 1.0 - nd * nd; /* nd_mnop(p2,p1,p2,p1) = nd_mnop(p4,p3,p4,p3) = 1 */
  if (den <= limit && den >= -limit) { *c1 = -DBL_MAX; *c2 = -DBL_MAX; return 1; }
  g_sub(p1, p3, p13);
  d2121 = g_dot(a, a); d4343 = g_dot(b, b); d4321 = g_dot(b, a); d1343 = g_dot(p13, b); d1321 = g_dot(p13, a);
  den = d2121 * d4343 - d4321 * d4321;
  *c1 = (d1343 * d4321 - d1321 * d4343) / den;
  *c2 = (d1343 + *c1 * d4321) / d4343;
  return 0;
}
/* project_points :292-333 */
__device__ static void g_project_points(const double *sundir, const double *origin, const double *normal, double *v1, double *v2, double *v3, double *v4) {
  const double big = 1e5;
  double *vs[4]; int q, m;
  vs[0] = v1; vs[1] = v2; vs[2] = v3; vs[3] = v4;
  if (fabs(g_dot(sundir, normal)) < sqrt(DBL_EPSILON)) {
    gv3 sp;
    for (m = 0; m < 3; ++m) sp[m] = -(fabs(normal[m]) - 1.0) * sundir[m];
    for (q = 0; q < 4; ++q) {
      double h = g_hit_plane(vs[q], normal, origin, normal);
      for (m = 0; m < 3; ++m) vs[q][m] = vs[q][m] + h * normal[m] - big * sp[m];
    }
  } else {
    for (q = 0; q < 4; ++q) {
      double h = g_hit_plane(vs[q], sundir, origin, normal);
      for (m = 0; m < 3; ++m) vs[q][m] = vs[q][m] + h * sundir[m];
    }
  }
}
/* rearange_projection :394-404 */
__device__ static void g_rearange_projection(const double *direction1, const double *origin2, const double *direction2,
                                  const double *origin3, const double *direction3, double *origin1) {
  double c21, c22, c31, c32, c;
  g_line_intersection(origin1, direction1, origin2, direction2, &c21, &c22);
  g_line_intersection(origin1, direction1, origin3, direction3, &c31, &c32);
  c = fmin(fmax(fmax(c21, c31), 0.0), 1.0);
  g_axpy(origin1, c, direction1, origin1);
}
/* rearange_projections :372-392 */
__device__ static void g_rearange_projections(const double *f1, const double *f2, const double *f3, const double *f4,
                                   double *v1, double *v2, double *v3, double *v4) {
  gv3 d, d2, d3;
  g_sub(f1, v1, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f4, f3, d2); g_sub(f3, f2, d3); g_rearange_projection(d, f3, d2, f2, d3, v1); }
  g_sub(f2, v2, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f4, f3, d2); g_sub(f1, f4, d3); g_rearange_projection(d, f3, d2, f4, d3, v2); }
  g_sub(f3, v3, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f1, f2, d2); g_sub(f1, f4, d3); g_rearange_projection(d, f2, d2, f4, d3, v3); }
  g_sub(f4, v4, d); if (g_norm(d) > DBL_EPSILON) { g_sub(f1, f2, d2); g_sub(f3, f2, d3); g_rearange_projection(d, f2, d2, f2, d3, v4); }
}
/* proj_var_to_edges :349-370 (the fourth point is moved along c3 - c2 although intersected along c4 - c1, as there) */
__device__ static void g_proj_var_to_edges(const double *c1, const double *c2, const double *c3, const double *c4, const double *v,
                                double *p1, double *p2, double *p3, double *p4) {
  gv3 d, e; double c, t;
  g_sub(c4, c1, d); g_sub(c4, c3, e); g_line_intersection(v, d, c3, e, &c, &t); g_axpy(v, c, d, p1);
  g_sub(c2, c1, d); g_sub(c3, c2, e); g_line_intersection(v, d, c2, e, &c, &t); g_axpy(v, c, d, p2);
  g_sub(c1, c2, d); g_sub(c1, c4, e); g_line_intersection(v, d, c4, e, &c, &t); g_axpy(v, c, d, p3);
  g_sub(c4, c1, d); g_sub(c1, c2, e); g_line_intersection(v, d, c2, e, &c, &t); g_sub(c3, c2, d); g_axpy(v, c, d, p4);
}
__device__ static double g_ext(double beta_s) { double x = fmax(beta_s, 10 * DBL_MIN); return expm1(-x) / (-x); } /* :276-283 */
__device__ static double g_num_dst(double s0, double l0, double h0, double c_ext, int n) { /* :248-274 */
  double dl, dh, ds, acc = 0.0; int i;
  if (n <= 0) n = 10;
  dl = l0 / n; dh = h0 / n; ds = s0 / n;
  for (i = 1; i <= n; ++i) { double j = i - 0.5; acc += dl * (j * dh) * g_ext(j * ds * c_ext); }
  return acc;
}
/* gomtrc_coeffs :135-246; slice is 0-based here */
__device__ static void g_gomtrc(const double *sundir, const double *f1, const double *f2, const double *f3, const double *f4, const double *f5,
                     const double *v1, const double *v2, const double *v3, const double *v4, double c_ext,
                     const int slice[3], double *coeffs, double sun_up_down, int nint) {
  const double small = sqrt(DBL_EPSILON);
  gv3 n0, normal, pl, pb, pt, pr, d, d2, hs;
  double s, cos_src_trgt, area1, area2, area3, aq, at, tot, sum;
  g_normal(f1, f2, f3, n0);
  s = g_hit_plane(f1, sundir, f5, n0); hs[0] = s * sundir[0]; hs[1] = s * sundir[1]; hs[2] = s * sundir[2]; s = g_norm(hs);
  g_normal(f1, f2, f5, normal);
  g_sub(f1, f2, d); g_sub(f1, f4, d2);
  cos_src_trgt = cos(acos(g_dot(d, d2) / (g_norm(d) * g_norm(d2))) - M_PI / 2);
#define G_S(v, f, nrm) do { double h_ = g_hit_plane(v, sundir, f, nrm); hs[0] = h_ * sundir[0]; hs[1] = h_ * sundir[1]; hs[2] = h_ * sundir[2]; s = g_norm(hs); } while (0)
#define G_LEN(a, b) (g_sub(a, b, d), g_norm(d))
  g_sub(v1, f1, d);
  if (g_norm(d) > small) {
    g_proj_var_to_edges(f1, f2, f3, f4, v1, pl, pb, pt, pr);
    area1 = g_quad(v1, pb, f3, pl) * exp(-c_ext * s);
    G_S(v1, f1, normal);
    aq = g_quad(v1, pr, f2, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f1, pr) * cos_src_trgt, G_LEN(pr, v1), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v1, f1, normal);
    aq = g_quad(v1, pl, f4, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f1, pt) * cos_src_trgt, G_LEN(pt, v1), c_ext, nint);
  } else if ((g_sub(v2, f2, d), g_norm(d)) > small) {
    g_proj_var_to_edges(f2, f1, f4, f3, v2, pl, pt, pb, pr);
    area1 = g_quad(v2, pl, f4, pt) * exp(-c_ext * s);
    G_S(v2, f2, normal);
    aq = g_quad(v2, pr, f1, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f2, pr) * cos_src_trgt, G_LEN(pr, v2), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v2, f2, normal);
    aq = g_quad(v2, pl, f3, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f2, pb) * cos_src_trgt, G_LEN(pb, v2), c_ext, nint);
  } else if ((g_sub(v3, f3, d), g_norm(d)) > small) {
    g_proj_var_to_edges(f4, f3, f2, f1, v3, pr, pb, pt, pl);
    area1 = g_quad(v3, pt, f1, pr) * exp(-c_ext * s);
    G_S(v3, f3, normal);
    aq = g_quad(v3, pl, f4, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f3, pl) * cos_src_trgt, G_LEN(pl, v3), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v3, f3, normal);
    aq = g_quad(v3, pr, f2, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f3, pb) * cos_src_trgt, G_LEN(pb, v3), c_ext, nint);
  } else if ((g_sub(v4, f4, d), g_norm(d)) > small) {
    g_proj_var_to_edges(f4, f3, f2, f1, v4, pr, pb, pt, pl);
    area1 = g_quad(v4, pr, f2, pb) * exp(-c_ext * s);
    G_S(v4, f4, normal);
    aq = g_quad(v4, pl, f3, pb) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f4, pl) * cos_src_trgt, G_LEN(pl, v4), c_ext, nint);
    area2 = aq + at;
    g_normal(f3, f2, f5, normal);
    G_S(v4, f4, normal);
    aq = g_quad(v4, pr, f1, pt) * g_ext(s * c_ext);
    at = g_num_dst(s, G_LEN(f4, pt) * cos_src_trgt, G_LEN(pt, v4), c_ext, nint);
  } else {
    area1 = g_quad(f1, f2, f3, f4) * exp(-c_ext * s);
    area2 = 0.0; aq = 0.0; at = 0.0;
  }
#undef G_S
#undef G_LEN
  area3 = aq + at;
  area1 = area1 - sun_up_down * area3;
  area3 = area3 + sun_up_down * area3;
  coeffs[slice[0]] = fmax(area1, 0.0); coeffs[slice[1]] = fmax(area2, 0.0); coeffs[slice[2]] = fmax(area3, 0.0);
  tot = g_quad(f1, f2, f3, f4);
  sum = coeffs[slice[0]] + coeffs[slice[1]] + coeffs[slice[2]];
  tot = fmax(tot, sum);
  coeffs[slice[0]] /= tot; coeffs[slice[1]] /= tot; coeffs[slice[2]] /= tot;
}
/* dir2dir3_geometric_coeffs :34-133.  Returns 1 if the extinction coefficient is too small (CHKERR there). */
__device__ int geom_dir2dir3_coeffs(const double *vertices, const double *sundir, double c_ext, int num_intervals, double *coeffs) {
  const double *a = vertices, *b = vertices + 3, *c = vertices + 6, *d = vertices + 9, *e = vertices + 12,
               *f = vertices + 15, *g = vertices + 18, *h = vertices + 21;
  const int sl_z[3] = {0, 6, 3}, sl_x[3] = {4, 7, 1}, sl_y[3] = {8, 5, 2};
  gv3 n, ms, p1, p2, p3, p4;
  double sun_up_down;
  if (c_ext < DBL_MIN) return 1;
  g_normal(e, f, h, n);
  ms[0] = -sundir[0]; ms[1] = -sundir[1]; ms[2] = -sundir[2];
  sun_up_down = fmin(g_dot(ms, n) >= 0.0 ? 1.0 : -1.0, 0.0);
  /* src z */
  g_copy(h, p1); g_copy(g, p2); g_copy(e, p3); g_copy(f, p4);
  g_normal(d, b, a, n);
  g_project_points(sundir, d, n, p1, p2, p3, p4);
  g_rearange_projections(d, c, a, b, p1, p2, p3, p4);
  g_gomtrc(sundir, d, c, a, b, g, p1, p2, p3, p4, c_ext, sl_z, coeffs, sun_up_down, num_intervals);
  /* src x */
  g_copy(h, p1); g_copy(d, p2); g_copy(b, p3); g_copy(f, p4);
  g_normal(c, a, e, n);
  g_project_points(sundir, a, n, p1, p2, p3, p4);
  g_rearange_projections(g, c, a, e, p1, p2, p3, p4);
  g_gomtrc(sundir, g, c, a, e, d, p1, p2, p3, p4, c_ext, sl_x, coeffs, sun_up_down, num_intervals);
  /* src y */
  g_copy(f, p1); g_copy(b, p2); g_copy(a, p3); g_copy(e, p4);
  g_normal(c, d, h, n);
  g_project_points(sundir, d, n, p1, p2, p3, p4);
  g_rearange_projections(h, d, c, g, p1, p2, p3, p4);
  g_gomtrc(sundir, h, d, c, g, b, p1, p2, p3, p4, c_ext, sl_y, coeffs, sun_up_down, num_intervals);
  return 0;
}

