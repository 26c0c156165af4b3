// Example 1: tetrahedral mesh (TetGen .node/.ele pair given by its stem), default upstream's cube fixture.
#include "example_common.h"

int main(int argc, char** argv) {
  return run_example(parse_example_options(argc, argv, "../src/test/test_data/CubeMesh_size256step16"), /*tet_mesh=*/true);
}
