// Shared body of the three known-answer programs (the counterparts of upstream's src/test/sanity2D.cc,
// sanity3D.cc and tetVol.cc): build the solver on a mesh fixture, swap the assembled operator for the
// MATLAB fixture, run solveFEM with the default parameters (= one V-cycle) and measure the distance to
// the stored answer.  Everything goes through the public FEMSolver surface a drop-in user sees.
#pragma once
#include <cmath>
#include <iostream>
#include <string>

#include "FEMSolver.h"

struct KnownAnswerCase {
  const char* mesh;      // fixture name below TEST_DATA_DIR (PLY file, or .node/.ele stem)
  bool tets;
  const char* matrix;    // sparse operator (.mat)
  const char* rhs;       // right-hand side (.mat)
  const char* answer;    // expected solution (.mat)
};

// returns ||x - x_answer||_2, or +inf when a fixture cannot be read
inline double known_answer_distance(const KnownAnswerCase& c) {
  const std::string dir = std::string(TEST_DATA_DIR) + "/";
  FEMSolver solver(dir + c.mesh, c.tets, /*verbose=*/true);
  if (solver.readMatlabSparseMatrix(dir + c.matrix) != 0) return HUGE_VAL;
  const size_t n = solver.getMatrixRows();
  Vector_h_CG rhs(n, 1.0), guess(n, 0.0), expected;
  if (solver.readMatlabArray(dir + c.rhs, &rhs) != 0) return HUGE_VAL;
  if (solver.readMatlabArray(dir + c.answer, &expected) != 0 || expected.size() != n) return HUGE_VAL;
  solver.solveFEM(&guess, &rhs);  // overwrites the guess with the solution
  double sq = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double d = guess[i] - expected[i];
    sq += d * d;
  }
  const double dist = std::sqrt(sq);
  std::cout << c.mesh << " (" << n << " rows): The error is " << dist << std::endl;
  return dist;
}
