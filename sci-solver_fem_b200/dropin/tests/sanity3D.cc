// 3-D egg-carton fixture (tet mesh simple.node/.ele with the simple*.mat system); upstream accepts a
// distance below 1 for this case.
#include "gtest/gtest.h"
#include "known_answer.h"

TEST(SanityTests, EggCarton3D) {
  const KnownAnswerCase egg3d = {"simple", true, "simple.mat", "simpleb.mat", "simpleAns.mat"};
  ASSERT_LT(known_answer_distance(egg3d), 1.0);
}
