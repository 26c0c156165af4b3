// Command line and driver shared by the two example programs.  They keep upstream's interface
//   ExampleN [-v] [-i mesh] [-A A.mat] [-b b.mat]
// (verbose, input mesh, optional operator and right-hand side in MATLAB v5 files) and add
//   --pcg  --tol t  --seed s
// because the reference has no flag for solverType_ = 1 (AMG-preconditioned CG) and seeds its
// aggregation from the wall clock.  Output files are upstream's: output.mat and <mesh name>.vtk.
#pragma once
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "FEMSolver.h"

struct ExampleOptions {
  std::string mesh, matrix, rhs;
  bool verbose = false, pcg = false;
  double tolerance = 1e-6;
  unsigned seed = 0;
};

inline ExampleOptions parse_example_options(int argc, char** argv, const std::string& default_mesh) {
  ExampleOptions o;
  o.mesh = default_mesh;
  auto value_of = [&](int& i) -> const char* { return (i + 1 < argc) ? argv[++i] : nullptr; };
  for (int i = 1; i < argc; ++i) {
    const std::string flag = argv[i];
    if (flag == "-v") o.verbose = true;
    else if (flag == "--pcg") o.pcg = true;
    else if (flag == "-i") { if (const char* v = value_of(i)) o.mesh = v; }
    else if (flag == "-A") { if (const char* v = value_of(i)) o.matrix = v; }
    else if (flag == "-b") { if (const char* v = value_of(i)) o.rhs = v; }
    else if (flag == "--tol") { if (const char* v = value_of(i)) o.tolerance = std::atof(v); }
    else if (flag == "--seed") { if (const char* v = value_of(i)) o.seed = (unsigned)std::atoi(v); }
  }
  return o;
}

// mesh path -> name of the .vtk file: directory removed, and for single-file formats (PLY) the extension too
inline std::string vtk_stem(const std::string& path, bool strip_extension) {
  const size_t slash = path.find_last_of("/\\");
  std::string stem = (slash == std::string::npos) ? path : path.substr(slash + 1);
  if (strip_extension) {
    const size_t dot = stem.find_last_of('.');
    if (dot != std::string::npos) stem.erase(dot);
  }
  return stem;
}

inline int run_example(const ExampleOptions& o, bool tet_mesh) {
  FEMSolver solver(o.mesh, tet_mesh, o.verbose);
  solver.seed_ = o.seed;
  if (o.pcg) {
    solver.solverType_ = 1;
    solver.tolerance_ = o.tolerance;
  }
  if (!o.matrix.empty() && solver.readMatlabSparseMatrix(o.matrix) != 0)
    std::cerr << "Failed to read in A matrix: " << o.matrix << std::endl;
  const size_t n = solver.getMatrixRows();
  Vector_h_CG rhs(n, 1.0), solution(n, 0.0);  // b = 1 unless a file is given; zero initial guess
  if (!o.rhs.empty() && solver.readMatlabArray(o.rhs, &rhs) != 0)
    std::cerr << "Failed to read in b array: " << o.rhs << std::endl;
  solver.solveFEM(&solution, &rhs);
  std::cout << "rows " << n << " iterations " << solver.iterations_ << " relres " << solver.relres_ << std::endl;
  if (solver.writeMatlabArray("output.mat", solution) != 0) std::cerr << "failed to write matlab file." << std::endl;
  solver.writeVTK(std::vector<double>(solution.begin(), solution.end()), vtk_stem(solver.filename_, !tet_mesh));
  return 0;
}
