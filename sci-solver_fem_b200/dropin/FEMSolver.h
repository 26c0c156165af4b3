// FEMSolver.h — the class a caller of SCI-Solver_FEM sees, backed by the B200 path.
//
// Source compatibility is the whole point of this header: constructor, methods, public data members and
// their defaults are those of the reference's class (src/FEMSolver.h:14-67, defaults src/FEMSolver.cu:9-44),
// so programs written against it — upstream's two examples and its three gtests — build unchanged.
// The members are grouped by what they control; C++ does not care about their order, callers only use names.
// Behind it: FEMSolver.cpp, a thin layer over the C ABI of libfemsolver_b200.so (include/femsolver_b200.h).
#ifndef __FEMSOLVER_H__
#define __FEMSOLVER_H__
#include <cmath>    // upstream callers use std::sqrt / strcmp through this header's transitive includes
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "TriMesh.h"
#include "tetmesh.h"
#include "types.h"

struct fsb_solver;  // opaque handle of the C ABI

class FEMSolver {
 public:
  // Reads <fname>.node/.ele (tets) or the PLY file <fname> (triangles) and assembles K + M on the GPU.
  FEMSolver(std::string fname = "../src/test/test_data/simple", bool isTetMesh = true, bool verbose = false);
  virtual ~FEMSolver();

  // ---- the solve path
  void getMatrixFromMesh();                                   // (re)assemble from the mesh
  void solveFEM(Vector_h_CG* x_h, Vector_h_CG* b_h);          // AMG setup + solve; *x_h: initial guess in, solution out
  size_t getMatrixRows();

  // ---- files either side of it (return 0 on success, like upstream)
  int readMatlabSparseMatrix(const std::string& filename);    // replaces the operator's values
  int readMatlabArray(const std::string& filename, Vector_h_CG* rhs);
  int writeMatlabArray(const std::string& filename, const Vector_h_CG& array);
  void writeVTK(std::vector<double> values, std::string fname);
  void checkMatrixForValidContents(Matrix_ell_h* A_h);

  // ---- input
  std::string filename_;
  TetMesh* tetMesh_;
  TriMesh* triMesh_;
  Matrix_ell_h A_h_;          // host copy of the operator (ELL, float values, as upstream keeps it)
  bool verbose_;
  int device_;

  // ---- AMG hierarchy
  int maxLevels_;             // 100
  int topSize_;               // 256: stop coarsening below this many rows
  int aggregatorType_;        // 0 = MIS based, 1 = METIS bottom-up
  int randMisParameters_;     // 90102: packed MIS depths and minimum aggregate size
  int partitionMaxSize_;      // 512
  double proOmega_;           // 0.67: prolongator smoothing

  // ---- smoother
  int preInnerIters_;         // 5
  int postInnerIters_;        // 5
  int postRelaxes_;           // 1
  double smootherWeight_;     // 1.0
  int dsType_;                // 0 (the only data structure upstream executes)

  // ---- outer iteration
  int solverType_;            // 0 = one V-cycle, 1 = AMG-preconditioned CG
  int maxIters_;              // 100
  double tolerance_;          // 1e-6, relative residual

  // ---- accepted and ignored, exactly as upstream ignores them
  int cycleIters_;
  int cycleType_;
  int convergeType_;
  int blockSize_;

  // ---- additions of this implementation
  unsigned seed_;             // aggregation seed (upstream: wall clock)
  int refLevel0NoPerm_;       // 1 = reproduce upstream's unpermuted level-0 vectors
  int iterations_;            // of the last solve
  double relres_;

 private:
  fsb_solver* impl_;
  bool A_from_file_;
  void pushParams();
  void pullMatrix();
};
#endif
