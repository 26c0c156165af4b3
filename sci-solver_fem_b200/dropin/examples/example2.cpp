// Example 2 (triangle mesh) of the reference (src/examples/example2.cu) against the drop-in FEMSolver:
//   Example1 [-v] [-i mesh_base] [-A A.mat] [-b b.mat] [--pcg] [--tol t] [--seed s]
// The first four flags are upstream's; --pcg/--tol/--seed are additive (solverType_=1 has no flag upstream).
#include <cstring>
#include "FEMSolver.h"

int main(int argc, char** argv) {
  std::string Aname = "", bName, fname = "../src/test/test_data/sphere_290verts.ply";
  bool verbose = false, pcg = false;
  double tol = 1e-6;
  unsigned seed = 0;
  for (int i = 0; i < argc; i++) {
    if (strcmp(argv[i], "-v") == 0) verbose = true;
    else if (strcmp(argv[i], "-i") == 0 && i + 1 < argc) fname = argv[++i];
    else if (strcmp(argv[i], "-b") == 0 && i + 1 < argc) bName = argv[++i];
    else if (strcmp(argv[i], "-A") == 0 && i + 1 < argc) Aname = argv[++i];
    else if (strcmp(argv[i], "--pcg") == 0) pcg = true;
    else if (strcmp(argv[i], "--tol") == 0 && i + 1 < argc) tol = atof(argv[++i]);
    else if (strcmp(argv[i], "--seed") == 0 && i + 1 < argc) seed = (unsigned)atoi(argv[++i]);
  }
  FEMSolver cfg(fname, false, verbose);
  if (pcg) { cfg.solverType_ = 1; cfg.tolerance_ = tol; }
  cfg.seed_ = seed;
  if (!Aname.empty() && cfg.readMatlabSparseMatrix(Aname) != 0) std::cerr << "Failed to read in A matrix: " << Aname << std::endl;
  Vector_h_CG b_h(cfg.getMatrixRows(), 1.0);
  if (!bName.empty() && cfg.readMatlabArray(bName, &b_h) != 0) std::cerr << "Failed to read in b array: " << bName << std::endl;
  Vector_h_CG x_h(cfg.getMatrixRows(), 0.0);
  cfg.solveFEM(&x_h, &b_h);
  std::cout << "rows " << cfg.getMatrixRows() << " iterations " << cfg.iterations_ << " relres " << cfg.relres_ << std::endl;
  if (cfg.writeMatlabArray("output.mat", x_h)) std::cerr << "failed to write matlab file." << std::endl;
  std::vector<double> vals(x_h.begin(), x_h.end());
  size_t pos = cfg.filename_.find_last_of("/");
  std::string outname = cfg.filename_.substr(pos == std::string::npos ? 0 : pos + 1);
  cfg.writeVTK(vals, outname);
  return 0;
}
