// 2-D egg-carton fixture (triangle mesh simple.ply with the simpleTri*.mat system); upstream accepts a
// distance below 100 for this case.
#include "gtest/gtest.h"
#include "known_answer.h"

TEST(SanityTests, EggCarton2D) {
  const KnownAnswerCase egg2d = {"simple.ply", false, "simpleTri.mat", "simpleTrib.mat", "simpleTriAns.mat"};
  ASSERT_LT(known_answer_distance(egg2d), 100.0);
}
