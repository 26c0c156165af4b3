// Volume-conductor fixture (tetVol.node/.ele with tetVolA/b/Ans.mat: a genuine FEM system whose pattern is
// the mesh pattern); upstream accepts a distance below 25 after its single V-cycle.
#include "gtest/gtest.h"
#include "known_answer.h"

TEST(SanityTests, TetVol) {
  const KnownAnswerCase tetvol = {"tetVol", true, "tetVolA.mat", "tetVolb.mat", "tetVolAns.mat"};
  ASSERT_LT(known_answer_distance(tetvol), 25.0);
}
