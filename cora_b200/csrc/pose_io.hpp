// pose_io.hpp -- host utilities either side of the solver: the odometry initialisation the
// reference's experiments start from, and export of a rounded solution to TUM / g2o text.
//
//   getOdomInitialization   examples/paper_experiments.cpp:426-534
//   saveSolnToTum / saveSolnToG20 / getRotation / getTranslation   src/CORA_utils.cpp:204-350
//
// Data model: the measurement stacks of cora_b200_assemble / cora_b200_pyfg_arrays (pose and landmark
// translations indexed 0..n-1, n..n+l-1; rot_R row-major d x d; rp_t the translation of every relative
// measurement) and N x r column-major solutions in the reference row order
// [d rows per pose | m range rows | n pose translations | l landmark translations].
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace cora_b200 {

// SplitMix64: the reference draws from Eigen's unseeded Random(); here every random draw is reproducible
struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return 2.0 * ((double)(next() >> 11) * (1.0 / 9007199254740992.0)) - 1.0; }  // U[-1, 1)
};

// x0 (N x rank, column-major) by composing the odometry chains.  Consecutive poses i -> i+1 joined by a
// pose-pose measurement form a chain (one chain per robot); the first chain starts at the identity, every further
// one at a random pose.  Landmarks: 10 * U[-1,1]^d.  Range rows: the normalised difference of the two
// translations -- reference_sign != 0 reproduces paper_experiments.cpp:500-506 (second - first), 0 uses the sign
// the data matrix implies (first - second: Q23 = D Omega_r A_r with A_r = -1 at the first id,
// src/CORA_problem.cpp:142-145,663).  The whole matrix is multiplied by a random SO(rank) element.
inline void odometry_initialization(int d, int64_t n, int64_t l, int64_t E, const int64_t *rp_i, const int64_t *rp_j,
                                    const double *rp_t, int64_t Ep, const int64_t *rot_i, const int64_t *rot_j,
                                    const double *rot_R, int64_t m, const int64_t *rg_a, const int64_t *rg_b, int rank,
                                    uint64_t seed, int reference_sign, double *X) {
  if (d != 2 && d != 3) throw std::invalid_argument("dimension must be 2 or 3");
  if (rank < d) throw std::invalid_argument("relaxation rank must be >= dim");
  const int64_t N = (int64_t)d * n + m + n + l, dn = (int64_t)d * n, T0 = dn + m;
  SplitMix rng(seed);
  std::vector<double> R((size_t)n * d * d, 0.0), t((size_t)(n + l) * d, 0.0);
  // odometry edge i -> i+1: rotation from the rot stack, translation from the rp stack
  std::map<std::pair<int64_t, int64_t>, int64_t> tr_of;
  for (int64_t k = 0; k < E; ++k)
    if (rp_j[k] == rp_i[k] + 1 && rp_j[k] < n) tr_of.emplace(std::make_pair(rp_i[k], rp_j[k]), k);
  std::vector<int64_t> next_rot((size_t)std::max<int64_t>(n, 1), -1);
  for (int64_t k = 0; k < Ep; ++k)
    if (rot_j[k] == rot_i[k] + 1 && next_rot[rot_i[k]] < 0) next_rot[rot_i[k]] = k;
  auto random_rotation = [&](double *Rm) {  // getRandomStartPose: a random proper rotation
    if (d == 2) {
      const double th = 3.141592653589793 * rng.uniform();
      Rm[0] = std::cos(th); Rm[1] = -std::sin(th); Rm[2] = std::sin(th); Rm[3] = std::cos(th);
    } else {
      double q[4], nq = 0.0;
      for (double &v : q) { v = rng.uniform(); nq += v * v; }
      nq = std::sqrt(nq > 0 ? nq : 1.0);
      const double x = q[0] / nq, y = q[1] / nq, z = q[2] / nq, w = q[3] / nq;
      const double M[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                           2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                           2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
      for (int e = 0; e < 9; ++e) Rm[e] = M[e];
    }
  };
  bool first_chain = true;
  for (int64_t i = 0; i < n; ++i) {
    const bool chained = i > 0 && next_rot[i - 1] >= 0 && tr_of.count(std::make_pair(i - 1, i)) > 0;
    double *Ri = &R[(size_t)i * d * d], *ti = &t[(size_t)i * d];
    if (!chained) {  // start of a chain
      if (first_chain) {
        for (int a = 0; a < d; ++a) Ri[a * d + a] = 1.0;
        first_chain = false;
      } else {
        random_rotation(Ri);
        for (int a = 0; a < d; ++a) ti[a] = 10.0 * rng.uniform();
      }
      continue;
    }
    const double *Rp = &R[(size_t)(i - 1) * d * d], *tp = &t[(size_t)(i - 1) * d];
    const double *Rm = rot_R + (size_t)next_rot[i - 1] * d * d;
    const double *tm = rp_t + (size_t)tr_of[std::make_pair(i - 1, i)] * d;
    for (int a = 0; a < d; ++a) {  // cur_pose = cur_pose * measurement
      double s = tp[a];
      for (int b = 0; b < d; ++b) s += Rp[a * d + b] * tm[b];
      ti[a] = s;
      for (int c = 0; c < d; ++c) {
        double v = 0.0;
        for (int b = 0; b < d; ++b) v += Rp[a * d + b] * Rm[b * d + c];
        Ri[a * d + c] = v;
      }
    }
  }
  for (int64_t j = 0; j < l; ++j)
    for (int a = 0; a < d; ++a) t[(size_t)(n + j) * d + a] = 10.0 * rng.uniform();
  std::vector<double> X0((size_t)N * rank, 0.0);  // column-major N x rank, data in the leading d columns
  auto at = [&](int64_t row, int c) -> double & { return X0[(size_t)c * N + row]; };
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < d; ++a)
      for (int c = 0; c < d; ++c) at(d * i + a, c) = R[(size_t)i * d * d + c * d + a];  // block = R_i^T
  for (int64_t x = 0; x < n + l; ++x)
    for (int c = 0; c < d; ++c) at(T0 + x, c) = t[(size_t)x * d + c];
  for (int64_t k = 0; k < m; ++k) {
    double diff[3] = {0, 0, 0}, nrm = 0.0;
    for (int c = 0; c < d; ++c) {
      const double a = t[(size_t)rg_a[k] * d + c], b = t[(size_t)rg_b[k] * d + c];
      diff[c] = reference_sign ? b - a : a - b;
      nrm += diff[c] * diff[c];
    }
    if (std::sqrt(nrm) < 1e-5) {
      nrm = 0.0;
      for (int c = 0; c < d; ++c) { diff[c] = rng.uniform(); nrm += diff[c] * diff[c]; }
    }
    nrm = std::sqrt(nrm);
    for (int c = 0; c < d; ++c) at(dn + k, c) = diff[c] / nrm;
  }
  // random SO(rank) factor: modified Gram-Schmidt of a random matrix, determinant fixed to +1
  std::vector<double> Qr((size_t)rank * rank);
  for (double &v : Qr) v = rng.uniform();
  for (int c = 0; c < rank; ++c) {
    for (int b = 0; b < c; ++b) {
      double s = 0.0;
      for (int a = 0; a < rank; ++a) s += Qr[(size_t)b * rank + a] * Qr[(size_t)c * rank + a];
      for (int a = 0; a < rank; ++a) Qr[(size_t)c * rank + a] -= s * Qr[(size_t)b * rank + a];
    }
    double s = 0.0;
    for (int a = 0; a < rank; ++a) s += Qr[(size_t)c * rank + a] * Qr[(size_t)c * rank + a];
    s = std::sqrt(s);
    if (!(s > 1e-12)) throw std::runtime_error("odometry initialisation: degenerate random rotation");
    for (int a = 0; a < rank; ++a) Qr[(size_t)c * rank + a] /= s;
  }
  {  // determinant by Gaussian elimination of a copy
    std::vector<double> A(Qr);
    double det = 1.0;
    for (int k = 0; k < rank; ++k) {
      int piv = k;
      for (int i = k + 1; i < rank; ++i)
        if (std::fabs(A[(size_t)k * rank + i]) > std::fabs(A[(size_t)k * rank + piv])) piv = i;
      if (piv != k) {
        for (int c = 0; c < rank; ++c) std::swap(A[(size_t)c * rank + k], A[(size_t)c * rank + piv]);
        det = -det;
      }
      det *= A[(size_t)k * rank + k];
      for (int i = k + 1; i < rank; ++i) {
        const double f = A[(size_t)k * rank + i] / A[(size_t)k * rank + k];
        for (int c = k; c < rank; ++c) A[(size_t)c * rank + i] -= f * A[(size_t)c * rank + k];
      }
    }
    if (det < 0)
      for (int a = 0; a < rank; ++a) Qr[(size_t)(rank - 1) * rank + a] = -Qr[(size_t)(rank - 1) * rank + a];
  }
  // X = X0 * Qr  (columns of Qr are orthonormal: Qr[c * rank + a] = Qr(a, c))
  for (int c = 0; c < rank; ++c)
    for (int64_t row = 0; row < N; ++row) {
      double s = 0.0;
      for (int a = 0; a < d; ++a) s += X0[(size_t)a * N + row] * Qr[(size_t)c * rank + a];
      X[(size_t)c * N + row] = s;
    }
}

// rotation of pose i from a rounded N x d solution (getRotation: the transpose of the stored block, checked)
inline void solution_rotation(int d, int64_t n, int64_t N, const double *X, int64_t i, double *Rm /* 3 x 3 */) {
  for (int e = 0; e < 9; ++e) Rm[e] = (e % 4 == 0) ? 1.0 : 0.0;
  for (int a = 0; a < d; ++a)
    for (int c = 0; c < d; ++c) Rm[c * 3 + a] = X[(size_t)c * N + (size_t)d * i + a];  // rot = block^T
  const double det = Rm[0] * (Rm[4] * Rm[8] - Rm[5] * Rm[7]) - Rm[1] * (Rm[3] * Rm[8] - Rm[5] * Rm[6]) +
                     Rm[2] * (Rm[3] * Rm[7] - Rm[4] * Rm[6]);
  if (std::fabs(det - 1.0) > 1e-6)
    throw std::runtime_error("Rotation matrix determinant is: " + std::to_string(det) + " not 1");
  double dev = 0.0;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += Rm[a * 3 + k] * Rm[b * 3 + k];
      s -= (a == b) ? 1.0 : 0.0;
      dev += s * s;
    }
  if (std::sqrt(dev) > 1e-6) throw std::runtime_error("Rotation matrix is not orthogonal");
}

// unit quaternion (x, y, z, w) of a rotation matrix, the branches of Eigen::Quaternion(Matrix3)
inline void rotation_to_quaternion(const double *Rm, double *q) {
  const double tr = Rm[0] + Rm[4] + Rm[8];
  if (tr > 0.0) {
    double s = std::sqrt(tr + 1.0);
    q[3] = 0.5 * s;
    s = 0.5 / s;
    q[0] = (Rm[7] - Rm[5]) * s; q[1] = (Rm[2] - Rm[6]) * s; q[2] = (Rm[3] - Rm[1]) * s;
  } else {
    int i = 0;
    if (Rm[4] > Rm[0]) i = 1;
    if (Rm[8] > Rm[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    double s = std::sqrt(Rm[i * 4] - Rm[j * 4] - Rm[k * 4] + 1.0);
    q[i] = 0.5 * s;
    s = 0.5 / s;
    q[3] = (Rm[k * 3 + j] - Rm[j * 3 + k]) * s;
    q[j] = (Rm[j * 3 + i] + Rm[i * 3 + j]) * s;
    q[k] = (Rm[k * 3 + i] + Rm[i * 3 + k]) * s;
  }
}

// poses first .. first+count-1 of a rounded solution as TUM ("time x y z qx qy qz qw") or g2o
// (VERTEX_SE3:QUAT / VERTEX_SE2) text, time = position in the list, default ostream formatting as the reference
inline void save_solution(const std::string &path, bool g2o, int d, int64_t n, int64_t m, int64_t nt, const double *X,
                          int64_t first, int64_t count) {
  if (d != 2 && d != 3) throw std::invalid_argument("dimension must be 2 or 3");
  if (first < 0 || count < 0 || first + count > n) throw std::invalid_argument("pose range outside the problem");
  const int64_t N = (int64_t)d * n + m + nt, T0 = (int64_t)d * n + m;
  std::ofstream out(path);
  if (!out.is_open()) throw std::runtime_error("Could not open file " + path);
  for (int64_t time = 0; time < count; ++time) {
    const int64_t i = first + time;
    double Rm[9], q[4];
    solution_rotation(d, n, N, X, i, Rm);
    const double x = X[(size_t)0 * N + T0 + i], y = X[(size_t)1 * N + T0 + i];
    const double z = d == 3 ? X[(size_t)2 * N + T0 + i] : 0.0;
    rotation_to_quaternion(Rm, q);
    if (!g2o) {
      out << time << " " << x << " " << y << " " << z << " " << q[0] << " " << q[1] << " " << q[2] << " " << q[3]
          << std::endl;
    } else if (d == 3) {
      out << "VERTEX_SE3:QUAT " << time << " " << x << " " << y << " " << z << " " << q[0] << " " << q[1] << " "
          << q[2] << " " << q[3] << "\n";
    } else {
      out << "VERTEX_SE2 " << time << " " << x << " " << y << " " << std::atan2(Rm[3], Rm[0]) << "\n";
    }
  }
}

}  // namespace cora_b200
