// nsb_kabsch.h — per-tet rotation of TetMesh::update_local_rotations (reference tet_mesh.cu:38-74), host and device.
//
// The reference forms the correlation matrix C = sum_k (orig_k - mean_orig)(def_k - mean_def)^T and takes R = U V^T from
// svd_eigen (editing/tools/svd3.h: McAdams et al., "Computing the SVD of 3x3 matrices with minimal branching"). That SVD keeps
// U and V proper rotations and lets the smallest singular value carry the sign of det(C), so R is ALWAYS a proper rotation —
// for an inverted tet it is the Kabsch rotation with the reflection fix, not the orthogonal polar factor (det -1), and for a
// flattened tet (sigma_3 = 0) it is still well defined. (Round 1 used a Newton polar iteration here: identical for ordinary tets,
// wrong for inverted / flat ones; tests/test_oracle_vs_ref.py::test_local_rotations_vs_reference_svd3 pins this to svd3.h.)
//
// Computed here in double: Jacobi eigen-decomposition of C^T C -> right-handed V (columns by descending eigenvalue), B = C V,
// U = Gram-Schmidt of (b1, b2) completed by u3 = u1 x u2, R = U V^T. Any sign / even-permutation ambiguity of V cancels in U V^T.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define NSB_HD __host__ __device__
#else
#define NSB_HD
#endif

namespace nsb {

// C: row-major 3x3 (row = canonical axis, column = deformed axis). R: column-major 3x3 floats. Identity when C has rank < 2.
NSB_HD inline void kabsch_rotation(const double C[3][3], float* R) {
	double S[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j) S[i][j] = C[0][i] * C[0][j] + C[1][i] * C[1][j] + C[2][i] * C[2][j];
	for (int sweep = 0; sweep < 12; ++sweep) {
		const double off = fabs(S[0][1]) + fabs(S[0][2]) + fabs(S[1][2]);
		if (off <= 1e-300 || off <= 1e-22 * (fabs(S[0][0]) + fabs(S[1][1]) + fabs(S[2][2]))) break;
		for (int p = 0; p < 2; ++p)
			for (int q = p + 1; q < 3; ++q) {
				if (fabs(S[p][q]) <= 1e-300) continue;
				const double theta = (S[q][q] - S[p][p]) / (2.0 * S[p][q]);
				const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
				const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
				for (int k = 0; k < 3; ++k) {  // S <- J^T S J, V <- V J
					const double skp = S[k][p], skq = S[k][q];
					S[k][p] = c * skp - s * skq;
					S[k][q] = s * skp + c * skq;
				}
				for (int k = 0; k < 3; ++k) {
					const double spk = S[p][k], sqk = S[q][k];
					S[p][k] = c * spk - s * sqk;
					S[q][k] = s * spk + c * sqk;
				}
				for (int k = 0; k < 3; ++k) {
					const double vkp = V[k][p], vkq = V[k][q];
					V[k][p] = c * vkp - s * vkq;
					V[k][q] = s * vkp + c * vkq;
				}
			}
	}
	int o0 = 0, o1 = 1, o2 = 2;  // eigenvalues descending
	if (S[o0][o0] < S[o1][o1]) { int x = o0; o0 = o1; o1 = x; }
	if (S[o0][o0] < S[o2][o2]) { int x = o0; o0 = o2; o2 = x; }
	if (S[o1][o1] < S[o2][o2]) { int x = o1; o1 = o2; o2 = x; }
	double v1[3] = {V[0][o0], V[1][o0], V[2][o0]}, v2[3] = {V[0][o1], V[1][o1], V[2][o1]};
	double v3[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
	double b1[3], b2[3];
	for (int r = 0; r < 3; ++r) {
		b1[r] = C[r][0] * v1[0] + C[r][1] * v1[1] + C[r][2] * v1[2];
		b2[r] = C[r][0] * v2[0] + C[r][1] * v2[1] + C[r][2] * v2[2];
	}
	const double n1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
	bool ok = n1 > 1e-150;
	double u1[3] = {1, 0, 0}, u2[3] = {0, 1, 0}, u3[3] = {0, 0, 1};
	if (ok) {
		for (int r = 0; r < 3; ++r) u1[r] = b1[r] / n1;
		const double d = u1[0] * b2[0] + u1[1] * b2[1] + u1[2] * b2[2];
		for (int r = 0; r < 3; ++r) u2[r] = b2[r] - d * u1[r];
		const double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
		ok = n2 > 1e-12 * n1;
		if (ok) {
			for (int r = 0; r < 3; ++r) u2[r] /= n2;
			u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
			u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
			u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
		}
	}
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c) R[c * 3 + r] = ok ? (float)(u1[r] * v1[c] + u2[r] * v2[c] + u3[r] * v3[c]) : (r == c ? 1.0f : 0.0f);
}

// correlation matrix of one tet + rotation
NSB_HD inline void tet_rotation(const float* verts, const float* orig, const uint32_t* tet, float* R) {
	double co[3] = {0, 0, 0}, cd[3] = {0, 0, 0};
	for (int j = 0; j < 4; ++j)
		for (int k = 0; k < 3; ++k) { co[k] += orig[3 * tet[j] + k]; cd[k] += verts[3 * tet[j] + k]; }
	for (int k = 0; k < 3; ++k) { co[k] /= 4.0; cd[k] /= 4.0; }
	double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
	for (int j = 0; j < 4; ++j)
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c) C[r][c] += (orig[3 * tet[j] + r] - co[r]) * (verts[3 * tet[j] + c] - cd[c]);
	kabsch_rotation(C, R);
}

}  // namespace nsb
