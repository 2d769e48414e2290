// nsb_host_geometry.cpp — the per-EDIT level of NeRFshop's two-level cage scheme (host code, like the
// reference): cage --MVC--> tet-mesh vertices --> per-cell tet lookup table + per-tet rotations.
// These run once per gizmo drag, not per sample; the per-sample level is device code (nsb_device.cuh).
//   nsb_build_tet_grid      <- TetMesh::build_tet_grid        (src/editing/datastructures/tet_mesh.cu:369-667)
//   nsb_compute_mvc         <- Cage::compute_mvc              (src/editing/datastructures/cage.cu:7-36); the
//                              coordinates are Ju/Schaefer/Warren 2005 "Mean Value Coordinates for Closed
//                              Triangular Meshes" (restated from the paper, the reference's mvc.h is GPL code)
//   nsb_interpolate_with_mvc<- Cage::interpolate_with_mvc     (cage.cu:39-55)
//   nsb_local_rotations     <- TetMesh::update_local_rotations (tet_mesh.cu:38-74)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nerfshop_b200.h"
#include "nsb_kabsch.h"

namespace {

struct P3 { float x, y, z; };
inline P3 p3(float x, float y, float z) { return P3{x, y, z}; }
inline P3 operator-(P3 a, P3 b) { return p3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline P3 operator+(P3 a, P3 b) { return p3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float dot(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline P3 cross(P3 a, P3 b) { return p3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline P3 ld(const float* v, uint32_t i) { return p3(v[3 * i], v[3 * i + 1], v[3 * i + 2]); }

// selection_utils.h:33-47
inline bool same_side(P3 v1, P3 v2, P3 v3, P3 v4, P3 p) {
	P3 n = cross(v2 - v1, v3 - v1);
	return std::signbit(dot(n, v4 - v1)) == std::signbit(dot(n, p - v1));
}
inline bool point_in_tet(P3 a, P3 b, P3 c, P3 d, P3 p) {
	return same_side(a, b, c, d, p) && same_side(b, c, d, a, p) && same_side(c, d, a, b, p) && same_side(d, a, b, c, p);
}

inline uint32_t expand_bits(uint32_t v) {
	v = (v * 0x00010001u) & 0xFF0000FFu;
	v = (v * 0x00000101u) & 0x0F00F00Fu;
	v = (v * 0x00000011u) & 0xC30C30C3u;
	v = (v * 0x00000005u) & 0x49249249u;
	return v;
}
inline uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) { return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2); }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (hi < v ? hi : v); }

// selection_utils.cu:65-83
inline P3 get_cell_pos(uint32_t x, uint32_t y, uint32_t z, uint32_t level) {
	float s = std::ldexp(1.0f, (int)level);
	return p3((((float)x + 0.5f) / 128.0f - 0.5f) * s + 0.5f, (((float)y + 0.5f) / 128.0f - 0.5f) * s + 0.5f, (((float)z + 0.5f) / 128.0f - 0.5f) * s + 0.5f);
}
inline void get_cell_at_pos(P3 pos, uint32_t level, int* out) {
	float s = std::ldexp(1.0f, -(int)level);
	float px = (pos.x - 0.5f) * s + 0.5f, py = (pos.y - 0.5f) * s + 0.5f, pz = (pos.z - 0.5f) * s + 0.5f;
	out[0] = clampi((int)(px * 128.0f), 0, 127);
	out[1] = clampi((int)(py * 128.0f), 0, 127);
	out[2] = clampi((int)(pz * 128.0f), 0, 127);
}

inline void project(const P3* pts, int n, P3 axis, float& mn, float& mx) {
	mn = INFINITY; mx = -INFINITY;
	for (int i = 0; i < n; ++i) {
		float v = dot(axis, pts[i]);
		if (v < mn) mn = v;
		if (v > mx) mx = v;
	}
}
// BoundingBox::intersects(Triangle) — separating axis test (bounding_box.cuh:126-178)
bool box_intersects_triangle(P3 bmin, P3 bmax, P3 a, P3 b, P3 c) {
	float tmin, tmax, bmn, bmx;
	const P3 bn[3] = {p3(1, 0, 0), p3(0, 1, 0), p3(0, 0, 1)};
	const float bminv[3] = {bmin.x, bmin.y, bmin.z}, bmaxv[3] = {bmax.x, bmax.y, bmax.z};
	P3 tri[3] = {a, b, c};
	for (int i = 0; i < 3; ++i) {
		project(tri, 3, bn[i], tmin, tmax);
		if (tmax < bminv[i] || tmin > bmaxv[i]) return false;
	}
	P3 verts[8] = {p3(bmin.x, bmin.y, bmin.z), p3(bmin.x, bmin.y, bmax.z), p3(bmin.x, bmax.y, bmin.z), p3(bmin.x, bmax.y, bmax.z),
	               p3(bmax.x, bmin.y, bmin.z), p3(bmax.x, bmin.y, bmax.z), p3(bmax.x, bmax.y, bmin.z), p3(bmax.x, bmax.y, bmax.z)};
	P3 n = cross(b - a, c - a);
	float len = std::sqrt(dot(n, n));
	if (len > 0.0f) n = p3(n.x / len, n.y / len, n.z / len);
	float off = dot(n, a);
	project(verts, 8, n, bmn, bmx);
	if (bmx < off || bmn > off) return false;
	const P3 edges[3] = {a - b, a - c, b - c};
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) {
			P3 axis = cross(edges[i], bn[j]);
			project(verts, 8, axis, bmn, bmx);
			project(tri, 3, axis, tmin, tmax);
			if (bmx < tmin || bmn > tmax) return false;
		}
	return true;
}

thread_local std::string g_err;

}  // namespace

// The library's nsb_last_error() lives in nsb_kernels.cu; geometry errors are plain argument errors.
extern "C" NsbStatus nsb_build_tet_grid(const float* vertices, uint32_t n_vertices, const uint32_t* tets, uint32_t n_tets, uint32_t* offsets, uint32_t* idx,
                                        uint64_t idx_capacity, uint64_t* n_idx, uint8_t* bitfield) {
	if (!vertices || !tets || !offsets || !n_idx) return NSB_ERR_INVALID;
	for (uint32_t i = 0; i < 4 * n_tets; ++i)
		if (tets[i] >= n_vertices) return NSB_ERR_INVALID;
	const uint32_t VOL = 128u * 128u * 128u;
	std::vector<std::pair<uint32_t, uint32_t>> marks;  // (cell, tet), generated in ascending tet order like the reference's merged thread lists
	if (bitfield) std::memset(bitfield, 0, NSB_BITFIELD_BYTES);
	const float corner[8][3] = {{-0.5f, -0.5f, -0.5f}, {-0.5f, -0.5f, 0.5f}, {-0.5f, 0.5f, -0.5f}, {0.5f, -0.5f, -0.5f},
	                            {0.5f, 0.5f, -0.5f},   {-0.5f, 0.5f, 0.5f},  {0.5f, -0.5f, 0.5f},  {0.5f, 0.5f, 0.5f}};
	for (uint32_t t = 0; t < n_tets; ++t) {
		P3 v[4] = {ld(vertices, tets[4 * t]), ld(vertices, tets[4 * t + 1]), ld(vertices, tets[4 * t + 2]), ld(vertices, tets[4 * t + 3])};
		P3 mn = v[0], mx = v[0];
		for (int j = 1; j < 4; ++j) {
			mn = p3(std::min(mn.x, v[j].x), std::min(mn.y, v[j].y), std::min(mn.z, v[j].z));
			mx = p3(std::max(mx.x, v[j].x), std::max(mx.y, v[j].y), std::max(mx.z, v[j].z));
		}
		for (uint32_t level = 0; level < NSB_NERF_CASCADES; ++level) {
			float scale = std::ldexp(1.0f, (int)level);
			int lo[3], hi[3];
			get_cell_at_pos(mn, level, lo);
			get_cell_at_pos(mx, level, hi);
			for (int x = lo[0]; x <= hi[0]; ++x)
				for (int y = lo[1]; y <= hi[1]; ++y)
					for (int z = lo[2]; z <= hi[2]; ++z) {
						P3 c = get_cell_pos((uint32_t)x, (uint32_t)y, (uint32_t)z, level);
						bool inside = false;
						for (int k = 0; k < 8 && !inside; ++k) {
							P3 q = p3(c.x + corner[k][0] * scale / 128.0f, c.y + corner[k][1] * scale / 128.0f, c.z + corner[k][2] * scale / 128.0f);
							inside = point_in_tet(v[0], v[1], v[2], v[3], q);
						}
						if (!inside) {
							float h = 0.5f * scale / 128.0f;
							P3 bmin = p3(c.x - h, c.y - h, c.z - h), bmax = p3(c.x + h, c.y + h, c.z + h);
							for (int j = 0; j < 4 && !inside; ++j) inside = box_intersects_triangle(bmin, bmax, v[j], v[(j + 1) % 4], v[(j + 2) % 4]);
						}
						if (inside) {
							uint32_t pos_idx = morton3D((uint32_t)x, (uint32_t)y, (uint32_t)z);
							uint32_t cell = level * VOL + pos_idx;
							marks.emplace_back(cell, t);
							if (bitfield) bitfield[pos_idx / 8 + (VOL * level) / 8] |= (uint8_t)(1u << (pos_idx % 8));
						}
					}
		}
	}
	// CSR (the reference counts in uint8_t, tet_mesh.cu:384; counts are kept in 32 bits here so a cell
	// listing >= 256 tets does not corrupt the offsets)
	std::vector<uint32_t> counts((size_t)NSB_GRID_CELLS, 0);
	for (auto& m : marks) counts[m.first]++;
	uint32_t running = 0;
	for (size_t i = 0; i < (size_t)NSB_GRID_CELLS; ++i) { offsets[i] = running; running += counts[i]; }
	offsets[NSB_GRID_CELLS] = running;
	*n_idx = marks.size();
	if (!idx || idx_capacity < marks.size()) return idx ? NSB_ERR_INVALID : NSB_OK;  // idx == NULL: size query
	std::fill(counts.begin(), counts.end(), 0);
	for (auto& m : marks) idx[offsets[m.first] + counts[m.first]++] = m.second;
	return NSB_OK;
}

// Mean value coordinates of `points` w.r.t. a closed triangle cage (Ju et al. 2005, Fig. 4), eps = 1e-8,
// special cases: point on a cage vertex -> one-hot; point on a triangle -> planar barycentric weights.
extern "C" NsbStatus nsb_compute_mvc(const float* cage_vertices, uint32_t n_cv, const uint32_t* cage_triangles, uint32_t n_tri, const float* points,
                                     uint32_t n_points, float gamma, float* weights) {
	if (!cage_vertices || !cage_triangles || !points || !weights) return NSB_ERR_INVALID;
	const double eps = 1e-8;
	std::vector<double> d(n_cv), w(n_cv);
	std::vector<double> ux(n_cv), uy(n_cv), uz(n_cv);
	for (uint32_t p = 0; p < n_points; ++p) {
		float* out = weights + (size_t)p * n_cv;
		for (uint32_t v = 0; v < n_cv; ++v) out[v] = 0.0f;
		const double ex = points[3 * p], ey = points[3 * p + 1], ez = points[3 * p + 2];
		bool done = false;
		for (uint32_t v = 0; v < n_cv && !done; ++v) {
			double dx = cage_vertices[3 * v] - ex, dy = cage_vertices[3 * v + 1] - ey, dz = cage_vertices[3 * v + 2] - ez;
			d[v] = std::sqrt(dx * dx + dy * dy + dz * dz);
			if (d[v] < eps) { out[v] = 1.0f; done = true; break; }
			ux[v] = dx / d[v]; uy[v] = dy / d[v]; uz[v] = dz / d[v];
		}
		if (done) continue;
		std::fill(w.begin(), w.end(), 0.0);
		double total = 0.0;
		for (uint32_t t = 0; t < n_tri && !done; ++t) {
			uint32_t id[3] = {cage_triangles[3 * t], cage_triangles[3 * t + 1], cage_triangles[3 * t + 2]};
			double l[3], theta[3], c[3], s[3];
			for (int i = 0; i < 3; ++i) {
				uint32_t a = id[(i + 1) % 3], b = id[(i + 2) % 3];
				double lx = ux[a] - ux[b], ly = uy[a] - uy[b], lz = uz[a] - uz[b];
				l[i] = std::sqrt(lx * lx + ly * ly + lz * lz);
				theta[i] = 2.0 * std::asin(std::min(1.0, l[i] / 2.0));
			}
			double h = (theta[0] + theta[1] + theta[2]) / 2.0;
			if (M_PI - h < eps) {  // the point lies on this triangle
				double wt[3];
				for (int i = 0; i < 3; ++i) wt[i] = std::sin(theta[i]) * l[(i + 2) % 3] * l[(i + 1) % 3];
				double sum = wt[0] + wt[1] + wt[2];
				for (uint32_t v = 0; v < n_cv; ++v) out[v] = 0.0f;
				for (int i = 0; i < 3; ++i) out[id[i]] = (float)(wt[i] / sum);
				done = true;
				break;
			}
			for (int i = 0; i < 3; ++i) c[i] = (2.0 * std::sin(h) * std::sin(h - theta[i])) / (std::sin(theta[(i + 1) % 3]) * std::sin(theta[(i + 2) % 3])) - 1.0;
			// orientation of (u0,u1,u2)
			double cx = uy[id[0]] * uz[id[1]] - uz[id[0]] * uy[id[1]];
			double cy = uz[id[0]] * ux[id[1]] - ux[id[0]] * uz[id[1]];
			double cz = ux[id[0]] * uy[id[1]] - uy[id[0]] * ux[id[1]];
			double sgn = (cx * ux[id[2]] + cy * uy[id[2]] + cz * uz[id[2]]) < 0.0 ? -1.0 : 1.0;
			bool skip = false;
			for (int i = 0; i < 3; ++i) {
				s[i] = sgn * std::sqrt(std::max(0.0, 1.0 - c[i] * c[i]));
				if (std::fabs(s[i]) < eps) skip = true;  // coplanar with the triangle but outside it
			}
			if (skip) continue;
			for (int i = 0; i < 3; ++i) {
				double wi = (theta[i] - c[(i + 1) % 3] * theta[(i + 2) % 3] - c[(i + 2) % 3] * theta[(i + 1) % 3]) / (2.0 * d[id[i]] * std::sin(theta[(i + 1) % 3]) * s[(i + 2) % 3]);
				w[id[i]] += wi;
				total += wi;
			}
		}
		if (done) continue;
		for (uint32_t v = 0; v < n_cv; ++v) out[v] = (float)(w[v] / total);
	}
	if (gamma > 1.0f) {  // cage.cu:25-35
		for (uint32_t p = 0; p < n_points; ++p) {
			float* out = weights + (size_t)p * n_cv;
			float sum = 0.0f;
			for (uint32_t v = 0; v < n_cv; ++v) { out[v] = std::pow(out[v], gamma); sum += out[v]; }
			for (uint32_t v = 0; v < n_cv; ++v) out[v] /= sum;
		}
	}
	return NSB_OK;
}

extern "C" NsbStatus nsb_interpolate_with_mvc(const float* weights, uint32_t n_points, uint32_t n_cv, const float* cage_vertices, float* points_out) {
	if (!weights || !cage_vertices || !points_out) return NSB_ERR_INVALID;
	for (uint32_t p = 0; p < n_points; ++p) {
		float acc[3] = {0, 0, 0};
		for (uint32_t v = 0; v < n_cv; ++v) {
			float wv = weights[(size_t)p * n_cv + v];
			acc[0] += wv * cage_vertices[3 * v];
			acc[1] += wv * cage_vertices[3 * v + 1];
			acc[2] += wv * cage_vertices[3 * v + 2];
		}
		points_out[3 * p] = acc[0]; points_out[3 * p + 1] = acc[1]; points_out[3 * p + 2] = acc[2];
	}
	return NSB_OK;
}

// Kabsch rotation per tet (always a proper rotation, like svd3.h's U V^T): nsb_kabsch.h
extern "C" NsbStatus nsb_local_rotations(const float* vertices, const float* original_vertices, const uint32_t* tets, uint32_t n_tets, float* rotations) {
	if (!vertices || !original_vertices || !tets || !rotations) return NSB_ERR_INVALID;
	for (uint32_t t = 0; t < n_tets; ++t) nsb::tet_rotation(vertices, original_vertices, tets + 4 * (size_t)t, rotations + 9 * (size_t)t);
	return NSB_OK;
}
