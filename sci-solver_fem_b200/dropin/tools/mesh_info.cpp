// mesh_info <file> : reads a triangle mesh through TriMesh::read (any supported format) and prints what a
// caller would see — counts, the first and last triangle, and order-sensitive checksums — one value per line.
// Host-only helper used by the CPU tests to compare the C++ readers with the Python ones.
#include <cstdio>

#include "FEMSolver.h"

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: mesh_info file\n"); return 2; }
  TriMesh* m = TriMesh::read(argv[1]);
  if (!m) return 1;
  double cs = 0.0;
  for (size_t i = 0; i < m->vertices.size(); i++)
    for (int j = 0; j < 3; j++) cs += (double)(i % 97 + 1) * (j + 1) * m->vertices[i][j];
  long long fs = 0;
  for (size_t i = 0; i < m->faces.size(); i++)
    for (int j = 0; j < 3; j++) fs += (long long)(i % 89 + 1) * (j + 1) * m->faces[i][j];
  std::printf("%zu\n%zu\n%.17g\n%lld\n", m->vertices.size(), m->faces.size(), cs, fs);
  if (!m->faces.empty()) {
    const TriMesh::Face &a = m->faces.front(), &b = m->faces.back();
    std::printf("%d %d %d\n%d %d %d\n", a[0], a[1], a[2], b[0], b[1], b[2]);
  }
  delete m;
  return 0;
}
