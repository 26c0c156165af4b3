// Example 2: triangle mesh (PLY file), default upstream's sphere fixture.
#include "example_common.h"

int main(int argc, char** argv) {
  return run_example(parse_example_options(argc, argv, "../src/test/test_data/sphere_290verts.ply"), /*tet_mesh=*/false);
}
