// pyfg.hpp -- PyFG text -> flattened measurement stacks (host).  Follows the reference's parser and
// data model: src/pyfg_text_parser.cpp:112-401 (13 line types, dimension from the first line,
// unknown keyword throws), include/CORA/Measurements.h:79-152 (precisions tau = d / tr(cov_t),
// kappa = 1/cov_theta (2D) or 3 / (2 tr(cov_R)) (3D), range precision 1/variance), src/CORA_problem.cpp:24-113
// (variables indexed in order of appearance, priors become factors from an auto-added origin pose
// "O0") -- with hash-set duplicate checks instead of the reference's O(M) std::find per added
// measurement.  Output order of the stacks: pose-pose, pose priors, pose-landmark,
// landmark priors (src/CORA_problem.cpp:190-294), i.e. the input of assemble_data_matrix().
#pragma once
#include <cmath>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace cora_b200 {

struct PyfgProblem {
  int d = 0;
  std::unordered_map<std::string, int64_t> pose_idx, landmark_idx;
  // translation-carrying factors in four groups, concatenated at the end
  struct RelT { std::string a, b; double t[3]; double tau; };
  struct RelR { std::string a, b; double R[9]; double kappa; };
  std::vector<RelT> pp_t, prior_t, pl_t, lprior_t;
  std::vector<RelR> pp_r, prior_r;
  struct Rng { std::string a, b; double r, w; };
  std::vector<Rng> ranges;
  std::unordered_set<std::string> pairs;
  bool has_priors = false;
  // flattened
  std::vector<int64_t> rp_i, rp_j, rot_i, rot_j, rg_a, rg_b;
  std::vector<double> rp_t, rp_tau, rot_R, rot_kappa, rg_r, rg_w;

  int64_t n() const { return (int64_t)pose_idx.size(); }
  int64_t l() const { return (int64_t)landmark_idx.size(); }

  void add_pose(const std::string &s) {  // src/CORA_problem.cpp:24-31
    if (pose_idx.count(s)) throw std::invalid_argument("Pose variable already exists");
    const int64_t k = (int64_t)pose_idx.size();
    pose_idx[s] = k;
  }
  void add_landmark(const std::string &s) {  // :33-40
    if (landmark_idx.count(s)) throw std::invalid_argument("Landmark variable already exists");
    const int64_t k = (int64_t)landmark_idx.size();
    landmark_idx[s] = k;
  }
  void check_pair(const char *kind, const std::string &a, const std::string &b, const char *msg) {
    const std::string key = std::string(kind) + "|" + (a <= b ? a + "|" + b : b + "|" + a);
    if (!pairs.insert(key).second) throw std::invalid_argument(msg);
  }
  void ensure_origin() {  // :80-86
    if (!has_priors) {
      has_priors = true;
      add_pose("O0");
    }
  }
  int64_t tr_idx(const std::string &s) const {  // translation index in [0, n + l), :998-1021
    auto it = pose_idx.find(s);
    if (it != pose_idx.end()) return it->second;
    auto jt = landmark_idx.find(s);
    if (jt != landmark_idx.end()) return n() + jt->second;
    throw std::invalid_argument("Unknown translation symbol: " + s);
  }
  int64_t rot_idx(const std::string &s) const {  // :964-974
    auto it = pose_idx.find(s);
    if (it == pose_idx.end()) throw std::invalid_argument("Unknown pose symbol: " + s);
    return it->second;
  }
  void flatten() {
    auto put_t = [&](const std::vector<RelT> &v) {
      for (const RelT &m : v) {
        rp_i.push_back(tr_idx(m.a)); rp_j.push_back(tr_idx(m.b));
        for (int k = 0; k < d; ++k) rp_t.push_back(m.t[k]);
        rp_tau.push_back(m.tau);
      }
    };
    auto put_r = [&](const std::vector<RelR> &v) {
      for (const RelR &m : v) {
        rot_i.push_back(rot_idx(m.a)); rot_j.push_back(rot_idx(m.b));
        for (int k = 0; k < d * d; ++k) rot_R.push_back(m.R[k]);
        rot_kappa.push_back(m.kappa);
      }
    };
    put_t(pp_t); put_t(prior_t); put_t(pl_t); put_t(lprior_t);
    put_r(pp_r); put_r(prior_r);
    for (const Rng &m : ranges) {
      rg_a.push_back(tr_idx(m.a)); rg_b.push_back(tr_idx(m.b));
      rg_r.push_back(m.r); rg_w.push_back(m.w);
    }
  }
};

namespace pyfg_detail {
inline void from_angle(double th, double *R) {  // :323-328, row-major 2 x 2
  const double c = std::cos(th), s = std::sin(th);
  R[0] = c; R[1] = -s; R[2] = s; R[3] = c;
}
inline void from_quat(double qx, double qy, double qz, double qw, double *R) {  // :330-338 (Eigen, no normalisation)
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
struct Tokens {
  std::vector<std::string> t;
  size_t pos = 0;
  const std::string &str() {
    if (pos >= t.size()) throw std::runtime_error("Unexpected end of line in PyFG file");
    return t[pos++];
  }
  double num() {
    const std::string &s = str();
    size_t used = 0;
    double v = 0;
    try { v = std::stod(s, &used); } catch (...) { used = 0; }
    if (used != s.size()) throw std::runtime_error("Could not parse number: " + s);
    return v;
  }
  // upper triangle row by row, :385-401; returns the diagonal only (all the precisions need)
  void symmetric_diag(int dim, double *diag) {
    for (int i = 0; i < dim; ++i)
      for (int j = i; j < dim; ++j) {
        const double v = num();
        if (i == j) diag[i] = v;
      }
  }
};
inline double trans_precision(const double *diag, int d) {  // Measurements.h:109-112,134-137
  double tr = 0;
  for (int k = 0; k < d; ++k) tr += diag[k];
  return (double)d / tr;
}
inline double rot_precision(const double *diag, int d) {  // Measurements.h:79-93
  if (d == 3) return 1.5 / (diag[3] + diag[4] + diag[5]);
  return 1.0 / diag[2];
}
}  // namespace pyfg_detail

inline void parse_pyfg(std::istream &in, PyfgProblem &P) {
  using namespace pyfg_detail;
  std::string line;
  bool first = true;
  while (std::getline(in, line)) {
    Tokens tk;
    {
      std::istringstream ls(line);
      std::string w;
      while (ls >> w) tk.t.push_back(w);
    }
    if (tk.t.empty()) throw std::runtime_error("Could not read item type from line " + line);
    const std::string kind = tk.str();
    if (first) {  // :41-97: the dimension comes from the first line
      if (kind == "VERTEX_SE2" || kind == "VERTEX_XY") P.d = 2;
      else if (kind == "VERTEX_SE3:QUAT" || kind == "VERTEX_XYZ") P.d = 3;
      else throw std::runtime_error("Could not determine dimension from first line " + line);
      first = false;
    }
    const int d = P.d;
    double diag[6];
    if (kind == "VERTEX_SE2" || kind == "VERTEX_SE3:QUAT") {  // ts sym <pose ignored>
      tk.str();
      P.add_pose(tk.str());
    } else if (kind == "VERTEX_XY" || kind == "VERTEX_XYZ") {  // sym <point ignored> (no timestamp)
      P.add_landmark(tk.str());
    } else if (kind == "EDGE_SE2" || kind == "EDGE_SE3:QUAT") {
      if ((kind == "EDGE_SE2") != (d == 2)) throw std::runtime_error("Edge type does not match the problem dimension");
      tk.str();
      PyfgProblem::RelT mt; PyfgProblem::RelR mr;
      mt.a = mr.a = tk.str(); mt.b = mr.b = tk.str();
      for (int k = 0; k < d; ++k) mt.t[k] = tk.num();
      if (d == 2) { from_angle(tk.num(), mr.R); }
      else { const double qx = tk.num(), qy = tk.num(), qz = tk.num(), qw = tk.num(); from_quat(qx, qy, qz, qw, mr.R); }
      tk.symmetric_diag(d == 2 ? 3 : 6, diag);
      mt.tau = trans_precision(diag, d); mr.kappa = rot_precision(diag, d);
      P.check_pair("p", mt.a, mt.b, "Relative pose measurement already exists");
      P.pp_t.push_back(mt); P.pp_r.push_back(mr);
    } else if (kind == "EDGE_SE2_XY" || kind == "EDGE_SE3_XYZ") {
      tk.str();
      PyfgProblem::RelT mt;
      mt.a = tk.str(); mt.b = tk.str();
      for (int k = 0; k < d; ++k) mt.t[k] = tk.num();
      tk.symmetric_diag(d, diag);
      mt.tau = trans_precision(diag, d);
      P.check_pair("pl", mt.a, mt.b, "Relative pose landmark measurement already exists");
      P.pl_t.push_back(mt);
    } else if (kind == "EDGE_RANGE") {
      tk.str();
      PyfgProblem::Rng m;
      m.a = tk.str(); m.b = tk.str();
      m.r = tk.num();
      m.w = 1.0 / tk.num();  // Measurements.h:151
      P.check_pair("r", m.a, m.b, "Range measurement already exists");
      P.ranges.push_back(m);
    } else if (kind == "VERTEX_SE2:PRIOR" || kind == "VERTEX_SE3:QUAT:PRIOR") {
      tk.str();
      PyfgProblem::RelT mt; PyfgProblem::RelR mr;
      mt.a = mr.a = "O0"; mt.b = mr.b = tk.str();
      for (int k = 0; k < d; ++k) mt.t[k] = tk.num();
      if (d == 2) { from_angle(tk.num(), mr.R); }
      else { const double qx = tk.num(), qy = tk.num(), qz = tk.num(), qw = tk.num(); from_quat(qx, qy, qz, qw, mr.R); }
      tk.symmetric_diag(d == 2 ? 3 : 6, diag);
      mt.tau = trans_precision(diag, d); mr.kappa = rot_precision(diag, d);
      P.check_pair("pp", mt.b, mt.b, "Pose prior already exists");
      P.prior_t.push_back(mt); P.prior_r.push_back(mr);
      P.ensure_origin();
    } else if (kind == "VERTEX_XY:PRIOR" || kind == "VERTEX_XYZ:PRIOR") {
      tk.str();
      PyfgProblem::RelT mt;
      mt.a = "O0"; mt.b = tk.str();
      for (int k = 0; k < d; ++k) mt.t[k] = tk.num();
      tk.symmetric_diag(d, diag);
      mt.tau = trans_precision(diag, d);
      P.check_pair("lp", mt.b, mt.b, "Landmark prior already exists");
      P.lprior_t.push_back(mt);
      P.ensure_origin();
    } else {
      throw std::runtime_error("Unknown item type " + kind);  // :157-159
    }
  }
  if (first) throw std::runtime_error("Could not read item type from line ");
  P.flatten();
}

}  // namespace cora_b200
