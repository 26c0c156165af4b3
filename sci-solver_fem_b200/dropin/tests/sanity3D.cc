// Restatement of the reference gtest src/test/sanity3D.cc against the drop-in FEMSolver (same inputs,
// same default parameters = one V-cycle, same assertion threshold).
#include "gtest/gtest.h"
#include "FEMSolver.h"
TEST(SanityTests, EggCarton3D) {
  FEMSolver cfg(std::string(TEST_DATA_DIR) + "/simple", true, true);
  cfg.readMatlabSparseMatrix(std::string(TEST_DATA_DIR) + "/simple.mat");
  Vector_h_CG b_h(cfg.getMatrixRows(), 1.0), x_h(cfg.getMatrixRows(), 0.), x_answer;
  cfg.readMatlabArray(std::string(TEST_DATA_DIR) + "/simpleb.mat", &b_h);
  cfg.solveFEM(&x_h, &b_h);
  cfg.readMatlabArray(std::string(TEST_DATA_DIR) + "/simpleAns.mat", &x_answer);
  double error = 0.f;
  for (size_t i = 0; i < cfg.getMatrixRows(); i++) error += (x_h[i] - x_answer[i]) * (x_h[i] - x_answer[i]);
  std::cout << "The error is : " << std::sqrt(error) << std::endl;
  ASSERT_TRUE(std::sqrt(error) < 1.);
}
