// FEMSolver.h — drop-in replacement of the reference's class FEMSolver (src/FEMSolver.h:14-67):
// same constructor, methods, public fields and defaults (src/FEMSolver.cu:9-44), so the upstream
// examples (src/examples/example{1,2}.cu) and gtests (src/test/*.cc) compile against it unchanged.
// Implementation: FEMSolver.cpp over the C-ABI of libfemsolver_b200.so (include/femsolver_b200.h).
// Additive surface: seed_, refLevel0NoPerm_, iterations_/relres_ of the last solve.
#ifndef __FEMSOLVER_H__
#define __FEMSOLVER_H__
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "TriMesh.h"
#include "tetmesh.h"
#include "types.h"

struct fsb_solver;

class FEMSolver {
 public:
  FEMSolver(std::string fname = "../src/test/test_data/simple", bool isTetMesh = true, bool verbose = false);
  virtual ~FEMSolver();
  void solveFEM(Vector_h_CG* x_h, Vector_h_CG* b_h);
  void getMatrixFromMesh();
  int readMatlabSparseMatrix(const std::string& filename);
  int readMatlabArray(const std::string& filename, Vector_h_CG* rhs);
  int writeMatlabArray(const std::string& filename, const Vector_h_CG& array);
  void checkMatrixForValidContents(Matrix_ell_h* A_h);
  void writeVTK(std::vector<double> values, std::string fname);
  size_t getMatrixRows();
  // data members (reference names, defaults and meaning)
  bool verbose_;
  std::string filename_;
  int maxLevels_;
  int maxIters_;
  int preInnerIters_;
  int postInnerIters_;
  int postRelaxes_;
  int cycleIters_;
  int dsType_;
  int topSize_;
  int randMisParameters_;
  int partitionMaxSize_;
  int aggregatorType_;
  int convergeType_;
  double tolerance_;
  int cycleType_;
  int solverType_;
  double smootherWeight_;
  double proOmega_;
  int device_;
  int blockSize_;
  TetMesh* tetMesh_;
  TriMesh* triMesh_;
  Matrix_ell_h A_h_;
  // additive
  unsigned seed_;
  int refLevel0NoPerm_;
  int iterations_;
  double relres_;

 private:
  fsb_solver* impl_;
  bool A_from_file_;
  void pushParams();
  void pullMatrix();
};
#endif
