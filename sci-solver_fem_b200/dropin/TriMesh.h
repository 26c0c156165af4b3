// tetmesh.h / TriMesh.h stand-ins: the data model the solve path consumes (vertices, tets/faces,
// matlabels, neighbors) and the input formats: TetGen .node/.ele; PLY (ascii / binary), OBJ, OFF, SM.
// Reference: src/core/include/tetmesh.h, src/core/cuda/tetmesh.cu:112-170 (need_neighbors),
// :251-376 (TetMesh::read); src/core/include/TriMesh.h, aggmis/cuda/TriMesh_connectivity.cu:96-131,
// TriMesh_io.cu:146-256 (format detection), :1239-1270 (tessellation).  Curvature / FIM members and the
// exotic formats (3DS, VVD, RAY, PLY strips) are out of scope.
#ifndef __FSB_MESHES_H__
#define __FSB_MESHES_H__
#include <string>
#include <vector>

template <int D, typename T>
struct Vec {
  T v[D];
  Vec() { for (int i = 0; i < D; i++) v[i] = T(); }
  T& operator[](int i) { return v[i]; }
  const T& operator[](int i) const { return v[i]; }
};
typedef Vec<3, double> point;

class TetMesh {
 public:
  struct Tet {
    int v[4];
    int& operator[](int i) { return v[i]; }
    const int& operator[](int i) const { return v[i]; }
  };
  std::vector<point> vertices;
  std::vector<Tet> tets;
  std::vector<int> matlabels;
  std::vector<std::vector<int> > neighbors;
  bool verbose = false;
  void set_verbose(bool v) { verbose = v; }
  void need_neighbors();
  void need_meshquality() {}  // upstream writes valance.txt / reratio.txt into the CWD; deliberately dropped
  static TetMesh* read(const char* nodefilename, const char* elefilename, const bool verbose = false);
};

class TriMesh {
 public:
  struct Face {
    int v[3];
    int& operator[](int i) { return v[i]; }
    const int& operator[](int i) const { return v[i]; }
  };
  static bool verbose;
  std::vector<point> vertices;
  std::vector<Face> faces;
  std::vector<std::vector<int> > neighbors;
  void set_verbose(bool v) { verbose = v; }
  void need_neighbors();
  void need_meshquality() {}
  static TriMesh* read(const char* filename);  // format by the first bytes; returns NULL on failure (TriMesh_io.cu:146-155)
};
#endif
