// FEMSolver.cpp — the reference's FEMSolver surface (src/FEMSolver.cu) over the C-ABI of
// libfemsolver_b200.so.  Host side only: parsing, format conversion, parameter forwarding; every
// numerical stage (pattern, assembly, AMG setup, PCG / V-cycle) runs in the CUDA library.
#include "FEMSolver.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <iterator>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <stdexcept>

#include "../../include/femsolver_b200.h"

bool TriMesh::verbose = false;

// ------------------------------------------------------------------------------------------- meshes
static bool next_data_line(std::ifstream& f, std::string& line) {
  while (f.good()) {
    std::getline(f, line);
    if (line.empty() || line.at(0) == '#') continue;
    return true;
  }
  return false;
}

// TetMesh::read, tetmesh.cu:251-376: %f into float, 0/1-based auto-detect, optional label column,
// second decrement when the minimum index is 1.
TetMesh* TetMesh::read(const char* nodefilename, const char* elefilename, const bool verb) {
  TetMesh* mesh = new TetMesh();
  mesh->set_verbose(verb);
  std::ifstream nodefile(nodefilename), elefile(elefilename);
  if (!nodefile.is_open() || !elefile.is_open()) {
    printf("node or ele file open failed!\n");
    exit(0);
  }
  std::string line;
  int nv = 0, tmp;
  if (!next_data_line(nodefile, line) || sscanf(line.c_str(), "%d %d %d %d", &nv, &tmp, &tmp, &tmp) != 4) {
    std::cerr << "Bad Node file" << std::endl;
    exit(0);
  }
  mesh->vertices.resize(nv);
  size_t i = 0;
  while (i < (size_t)nv && next_data_line(nodefile, line)) {
    float x, y, z;
    if (sscanf(line.c_str(), "%d %f %f %f", &tmp, &x, &y, &z) != 4) {
      std::cerr << "Bad Node file, line # " << i << std::endl;
      exit(0);
    }
    mesh->vertices[i][0] = x; mesh->vertices[i][1] = y; mesh->vertices[i][2] = z;
    i++;
  }
  int ne = 0, haslabel = 0;
  if (!next_data_line(elefile, line) || sscanf(line.c_str(), "%d %d %d", &ne, &tmp, &haslabel) != 3) {
    std::cerr << "Bad Ele file" << std::endl;
    exit(0);
  }
  mesh->tets.resize(ne);
  mesh->matlabels.resize(ne, 0);
  bool zero_based = false;
  i = 0;
  while (i < (size_t)ne && next_data_line(elefile, line)) {
    int t[4], mat = 0;
    int got = haslabel == 0 ? sscanf(line.c_str(), "%d %d %d %d %d", &tmp, &t[0], &t[1], &t[2], &t[3])
                            : sscanf(line.c_str(), "%d %d %d %d %d %d", &tmp, &t[0], &t[1], &t[2], &t[3], &mat);
    if (got != (haslabel == 0 ? 5 : 6)) {
      std::cerr << "Bad Ele file, line # " << i << std::endl;
      exit(0);
    }
    if (haslabel != 0) mesh->matlabels[i] = mat;
    for (int j = 0; j < 4; j++) { mesh->tets[i][j] = t[j]; if (t[j] == 0) zero_based = true; }
    i++;
  }
  if (!zero_based) for (auto& t : mesh->tets) for (int j = 0; j < 4; j++) t[j]--;
  int minidx = INT32_MAX;
  for (auto& t : mesh->tets) for (int j = 0; j < 4; j++) minidx = std::min(minidx, t[j]);
  if (minidx == 1) for (auto& t : mesh->tets) for (int j = 0; j < 4; j++) t[j]--;
  return mesh;
}

// ASCII PLY with x y z first (TriMesh_io.cu:259, :874-878 use %lf)
// ---- triangle-mesh input.  TriMesh::read upstream recognises the format from the first bytes of the file
// (TriMesh_io.cu:160-256).  Covered: PLY (ascii, binary little/big endian; any scalar property types, the
// vertex_indices list with any integer count/index types, faces as `element face` or as `element tristrips`, other
// elements — also range_grid, whose upstream reader is commented out — skipped), 3DS, VVD, RAY, OBJ, OFF, old-style SM.
// Polygons are cut into triangles by upstream's rule (:1239-1270).
// Deviations where upstream cannot work: its 3DS loop returns false at the end of every well-formed file (feof is only set
// by the failing read, :508-513); its RAY reader scans "%f" into doubles (:634, undefined behaviour); it reads triangle
// strips but never unpacks them (convert_strips / need_faces commented out, :478) — here all three deliver the mesh.
namespace {

struct PlyProp { bool list; std::string name; int size, csize; char kind, ckind; };  // kind: 'i' signed, 'u' unsigned, 'f' float
struct PlyElem { std::string name; long count; std::vector<PlyProp> props; };

bool ply_type(const std::string& t, int& size, char& kind) {
  static const struct { const char* n; int s; char k; } T[] = {
      {"char", 1, 'i'}, {"int8", 1, 'i'}, {"uchar", 1, 'u'}, {"uint8", 1, 'u'}, {"short", 2, 'i'}, {"int16", 2, 'i'},
      {"ushort", 2, 'u'}, {"uint16", 2, 'u'}, {"int", 4, 'i'}, {"int32", 4, 'i'}, {"uint", 4, 'u'}, {"uint32", 4, 'u'},
      {"float", 4, 'f'}, {"float32", 4, 'f'}, {"double", 8, 'f'}, {"float64", 8, 'f'}};
  for (const auto& e : T)
    if (t == e.n) { size = e.s; kind = e.k; return true; }
  return false;
}

// one binary scalar of the given type at p (byte-swapped when the file's endianness is not the host's)
double bin_scalar(const unsigned char* p, int size, char kind, bool swap) {
  unsigned char b[8];
  for (int i = 0; i < size; i++) b[i] = p[swap ? size - 1 - i : i];
  if (kind == 'f') {
    if (size == 4) { float v; std::memcpy(&v, b, 4); return v; }
    double v; std::memcpy(&v, b, 8); return v;
  }
  if (size == 1) return kind == 'i' ? (double)(signed char)b[0] : (double)b[0];
  if (size == 2) { if (kind == 'i') { short v; std::memcpy(&v, b, 2); return v; } unsigned short v; std::memcpy(&v, b, 2); return v; }
  if (kind == 'i') { int v; std::memcpy(&v, b, 4); return v; }
  unsigned v; std::memcpy(&v, b, 4); return v;
}

void tessellate(const std::vector<point>& verts, const std::vector<int>& poly, std::vector<TriMesh::Face>& out) {
  auto push = [&](int a, int b, int c) { TriMesh::Face f; f[0] = a; f[1] = b; f[2] = c; out.push_back(f); };
  const size_t k = poly.size();
  if (k < 3) return;
  if (k == 3) { push(poly[0], poly[1], poly[2]); return; }
  if (k == 4) {  // along the shorter diagonal
    auto d2 = [&](int a, int b) { double s = 0; for (int j = 0; j < 3; j++) { double d = verts[a][j] - verts[b][j]; s += d * d; } return s; };
    const int i = d2(poly[0], poly[2]) < d2(poly[1], poly[3]) ? 0 : 1;
    push(poly[i], poly[(i + 1) % 4], poly[(i + 2) % 4]);
    push(poly[i], poly[(i + 2) % 4], poly[(i + 3) % 4]);
    return;
  }
  for (size_t i = 2; i < k; i++) push(poly[0], poly[i - 1], poly[i]);
}

bool finish(TriMesh* m, const std::vector<std::vector<int> >& polys) {
  const int nv = (int)m->vertices.size();
  for (const auto& p : polys) {
    for (int i : p)
      if (i < 0 || i >= nv) return false;
    tessellate(m->vertices, p, m->faces);
  }
  return !(m->vertices.empty() && m->faces.empty());
}

// strips separated by -1 -> triangles, orientation flipped on every second triangle; degenerate ones only stitch strips
void unpack_tstrips(const std::vector<int>& idx, std::vector<std::vector<int> >& polys) {
  std::vector<int> run;
  auto flush = [&]() {
    for (size_t i = 0; i + 2 < run.size(); i++) {
      const int a = run[i], b = run[i + 1], c = run[i + 2];
      if (a == b || b == c || a == c) continue;
      if (i % 2 == 0) polys.push_back({a, b, c}); else polys.push_back({b, a, c});
    }
    run.clear();
  };
  for (int v : idx) { if (v < 0) flush(); else run.push_back(v); }
  flush();
}

bool parse_ply(const std::string& data, TriMesh* m) {
  const size_t end = data.find("end_header");
  if (end == std::string::npos) return false;
  const size_t body0 = data.find('\n', end);
  if (body0 == std::string::npos) return false;
  std::istringstream hs(data.substr(0, end));
  std::string line, fmt;
  std::vector<PlyElem> elems;
  while (std::getline(hs, line)) {
    std::istringstream ls(line);
    std::string w;
    if (!(ls >> w)) continue;
    if (w == "format") ls >> fmt;
    else if (w == "element") { PlyElem e; ls >> e.name >> e.count; elems.push_back(e); }
    else if (w == "property" && !elems.empty()) {
      PlyProp pr;
      std::string t;
      ls >> t;
      pr.list = (t == "list");
      if (pr.list) {
        std::string ct, it;
        ls >> ct >> it >> pr.name;
        if (!ply_type(ct, pr.csize, pr.ckind) || !ply_type(it, pr.size, pr.kind)) return false;
      } else {
        ls >> pr.name;
        pr.csize = 0; pr.ckind = 'u';
        if (!ply_type(t, pr.size, pr.kind)) return false;
      }
      elems.back().props.push_back(pr);
    }
  }
  const bool ascii = (fmt == "ascii");
  if (!ascii && fmt != "binary_little_endian" && fmt != "binary_big_endian") return false;
  const unsigned short probe = 1;
  const bool host_le = *reinterpret_cast<const unsigned char*>(&probe) == 1;
  const bool swap = !ascii && ((fmt == "binary_little_endian") != host_le);
  std::vector<std::vector<int> > polys;
  std::istringstream as;
  if (ascii) as.str(data.substr(body0 + 1));
  const unsigned char* p = reinterpret_cast<const unsigned char*>(data.data()) + body0 + 1;
  const unsigned char* pend = reinterpret_cast<const unsigned char*>(data.data()) + data.size();
  auto next = [&](int size, char kind, double& v) -> bool {
    if (ascii) return (bool)(as >> v);
    if (p + size > pend) return false;
    v = bin_scalar(p, size, kind, swap);
    p += size;
    return true;
  };
  for (const PlyElem& e : elems) {
    const bool is_vertex = (e.name == "vertex"), is_face = (e.name == "face");
    if (is_vertex) m->vertices.resize(e.count);
    for (long r = 0; r < e.count; r++) {
      for (const PlyProp& pr : e.props) {
        double v;
        if (!pr.list) {
          if (!next(pr.size, pr.kind, v)) return false;
          if (is_vertex) {
            if (pr.name == "x") m->vertices[r][0] = v;
            else if (pr.name == "y") m->vertices[r][1] = v;
            else if (pr.name == "z") m->vertices[r][2] = v;
          }
        } else {
          if (!next(pr.csize, pr.ckind, v)) return false;
          const long k = (long)v;
          std::vector<int> poly;
          for (long j = 0; j < k; j++) {
            if (!next(pr.size, pr.kind, v)) return false;
            poly.push_back((int)v);
          }
          if (is_face && (pr.name == "vertex_indices" || pr.name == "vertex_index")) polys.push_back(poly);
          if (e.name == "tristrips" && pr.name == "vertex_indices") unpack_tstrips(poly, polys);
        }
      }
    }
  }
  return finish(m, polys);
}

// whitespace-separated tokens of a text file with '#' comments removed
std::vector<std::string> text_tokens(const std::string& data) {
  std::vector<std::string> toks;
  std::istringstream ds(data);
  std::string line, w;
  while (std::getline(ds, line)) {
    const size_t h = line.find('#');
    if (h != std::string::npos) line.erase(h);
    std::istringstream ls(line);
    while (ls >> w) toks.push_back(w);
  }
  return toks;
}

bool parse_obj(const std::string& data, TriMesh* m) {
  std::vector<std::vector<int> > polys;
  std::istringstream ds(data);
  std::string line;
  while (std::getline(ds, line)) {
    std::istringstream ls(line);
    std::string w;
    if (!(ls >> w) || w[0] == '#') continue;
    if (w == "v") {
      point q;
      if (!(ls >> q[0] >> q[1] >> q[2])) return false;
      m->vertices.push_back(q);
    } else if (w == "f" || w == "t") {
      std::vector<int> poly;
      while (ls >> w) {  // a/b/c groups: the first integer is the vertex; 1-based, negative = relative
        char* endp = NULL;
        const long i = std::strtol(w.c_str(), &endp, 10);
        if (endp == w.c_str()) break;
        poly.push_back(i < 0 ? (int)(i + (long)m->vertices.size()) : (int)(i - 1));
      }
      polys.push_back(poly);
    }
  }
  return finish(m, polys);
}

bool parse_counted(const std::vector<std::string>& t, size_t pos, bool off_style, TriMesh* m) {
  // OFF: nverts nfaces [nedges] | vertices | k i0..ik-1 per face.   SM: nverts | vertices | nfaces | 3 indices per face
  auto num = [&](size_t i, double& v) { if (i >= t.size()) return false; char* e = NULL; v = std::strtod(t[i].c_str(), &e); return e != t[i].c_str(); };
  double v;
  if (!num(pos++, v)) return false;
  const long nv = (long)v;
  long nf = 0;
  if (off_style) { if (!num(pos++, v)) return false; nf = (long)v; pos++; }
  m->vertices.resize(nv);
  for (long i = 0; i < nv; i++)
    for (int j = 0; j < 3; j++) { if (!num(pos++, v)) return false; m->vertices[i][j] = v; }
  if (!off_style) { if (!num(pos++, v)) return finish(m, {}); nf = (long)v; }
  std::vector<std::vector<int> > polys;
  for (long f = 0; f < nf; f++) {
    long k = 3;
    if (off_style) { if (!num(pos++, v)) return false; k = (long)v; }
    std::vector<int> poly;
    for (long j = 0; j < k; j++) { if (!num(pos++, v)) return false; poly.push_back((int)v); }
    polys.push_back(poly);
  }
  return finish(m, polys);
}

template <typename T>
bool get_bin(const std::string& d, size_t& pos, bool big_endian, T& v) {
  if (pos + sizeof(T) > d.size()) return false;
  unsigned char b[sizeof(T)];
  const unsigned short probe = 1;
  const bool host_le = *reinterpret_cast<const unsigned char*>(&probe) == 1;
  for (size_t i = 0; i < sizeof(T); i++) b[i] = (unsigned char)d[pos + ((big_endian == host_le) ? sizeof(T) - 1 - i : i)];
  std::memcpy(&v, b, sizeof(T));
  pos += sizeof(T);
  return true;
}

// 3D Studio chunks (TriMesh_io.cu:503-576)
bool parse_3ds(const std::string& d, TriMesh* m) {
  std::vector<std::vector<int> > polys;
  size_t pos = 0;
  int mstart = 0;
  while (pos + 6 <= d.size()) {
    unsigned short id; unsigned len;
    get_bin(d, pos, false, id); get_bin(d, pos, false, len);
    if (id == 0x4d4d || id == 0x3d3d) continue;                       // entered
    if (id == 0x4000) {                                               // object: zero-terminated name, then entered
      const size_t e = d.find('\0', pos);
      if (e == std::string::npos) return false;
      pos = e + 1;
    } else if (id == 0x4100) mstart = (int)m->vertices.size();
    else if (id == 0x4110) {
      unsigned short nv;
      if (!get_bin(d, pos, false, nv)) return false;
      for (int i = 0; i < nv; i++) {
        float x, y, z;
        if (!get_bin(d, pos, false, x) || !get_bin(d, pos, false, y) || !get_bin(d, pos, false, z)) return false;
        point q; q[0] = x; q[1] = y; q[2] = z;
        m->vertices.push_back(q);
      }
    } else if (id == 0x4120) {
      unsigned short nf;
      if (!get_bin(d, pos, false, nf)) return false;
      for (int i = 0; i < nf; i++) {
        unsigned short a, b, c, flags;
        if (!get_bin(d, pos, false, a) || !get_bin(d, pos, false, b) || !get_bin(d, pos, false, c) || !get_bin(d, pos, false, flags)) return false;
        polys.push_back({mstart + a, mstart + b, mstart + c});
      }
    } else pos += len >= 6 ? len - 6 : 0;
  }
  return finish(m, polys);
}

// VIVID range scans, big-endian (TriMesh_io.cu:580-621)
bool parse_vvd(const std::string& d, TriMesh* m) {
  size_t pos = 5 + 127;
  int nv, nf;
  if (!get_bin(d, pos, true, nv) || nv < 0) { std::cerr << "Couldn't read vertex count" << std::endl; return false; }
  for (int i = 0; i < nv; i++) {
    double x, y, z;
    if (!get_bin(d, pos, true, x) || !get_bin(d, pos, true, y) || !get_bin(d, pos, true, z)) { std::cerr << "Couldn't read vertex" << std::endl; return false; }
    point q; q[0] = x; q[1] = y; q[2] = z;
    m->vertices.push_back(q);
  }
  if (!get_bin(d, pos, true, nf)) { std::cerr << "Couldn't read face count" << std::endl; return false; }
  std::vector<std::vector<int> > polys;
  for (int f = 0; f < nf; f++) {
    int k;
    if (!get_bin(d, pos, true, k) || k < 0) return false;
    std::vector<int> poly(k);
    for (int j = 0; j < k; j++) if (!get_bin(d, pos, true, poly[j])) return false;
    polys.push_back(poly);
  }
  return finish(m, polys);
}

// ray-tracer scene text (TriMesh_io.cu:625-647)
bool parse_ray(const std::string& d, TriMesh* m) {
  std::istringstream ds(d);
  std::vector<std::string> t;
  std::string w;
  while (ds >> w) t.push_back(w);
  std::vector<std::vector<int> > polys;
  for (size_t i = 0; i < t.size();) {
    if (t[i].compare(0, 7, "#vertex") == 0 && i + 3 < t.size()) {
      point q;
      for (int j = 0; j < 3; j++) q[j] = std::strtod(t[i + 1 + j].c_str(), NULL);
      m->vertices.push_back(q);
      i += 4;
    } else if (t[i].compare(0, 15, "#shape_triangle") == 0 && i + 4 < t.size()) {
      polys.push_back({std::atoi(t[i + 2].c_str()), std::atoi(t[i + 3].c_str()), std::atoi(t[i + 4].c_str())});
      i += 5;
    } else i++;
  }
  return finish(m, polys);
}

bool is_ray(const std::string& d) {  // the word after '#': material / vertex / shape_... (TriMesh_io.cu:211-219)
  std::istringstream ds(d.substr(1, 64));
  std::string w;
  if (!(ds >> w)) return false;
  return w.compare(0, 8, "material") == 0 || w.compare(0, 6, "vertex") == 0 || w.compare(0, 6, "shape_") == 0;
}

}  // namespace

TriMesh* TriMesh::read(const char* filename) {
  if (!filename || !*filename) return NULL;
  std::ifstream f(filename, std::ios::binary);
  if (!f.is_open()) { perror("fopen"); return NULL; }
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  TriMesh* m = new TriMesh();
  bool ok = false;
  if (data.empty()) std::cerr << "Can't read header" << std::endl;
  else if (data.compare(0, 3, "ply") == 0) ok = parse_ply(data, m);
  else if (data.compare(0, 2, "MM") == 0) ok = parse_3ds(data, m);
  else if (data.compare(0, 5, "VIVID") == 0) ok = parse_vvd(data, m);
  else if (data[0] == '#' && is_ray(data)) ok = parse_ray(data, m);
  else if (data.compare(0, 3, "OFF") == 0) { std::vector<std::string> t = text_tokens(data); ok = parse_counted(t, 1, true, m); }
  else if (std::strchr("#vufgso", data[0])) ok = parse_obj(data, m);
  else if (std::isdigit((unsigned char)data[0])) ok = parse_counted(text_tokens(data), 0, false, m);
  else std::cerr << "Unknown file type" << std::endl;
  if (!ok) {
    std::cerr << "\nError reading file [" << filename << "]" << std::endl;
    delete m;
    return NULL;
  }
  return m;
}

// neighbours in discovery order (tetmesh.cu:112-170 / TriMesh_connectivity.cu:96-131).  The solver
// itself takes adjacency from the device-built pattern; these exist for callers that read
// `mesh->neighbors` directly.
template <typename Mesh, typename Elems>
static void neighbors_by_discovery(Mesh* m, const Elems& elems, int npe) {
  if (!m->neighbors.empty()) return;
  m->neighbors.resize(m->vertices.size());
  for (size_t e = 0; e < elems.size(); e++)
    for (int j = 0; j < npe; j++) {
      std::vector<int>& me = m->neighbors[elems[e][j]];
      for (int d = 1; d < npe; d++) {
        int nb = elems[e][(j + d) % npe];
        if (std::find(me.begin(), me.end(), nb) == me.end()) me.push_back(nb);
      }
    }
}
void TetMesh::need_neighbors() { neighbors_by_discovery(this, tets, 4); }
void TriMesh::need_neighbors() { neighbors_by_discovery(this, faces, 3); }

// ------------------------------------------------------------------------------------------- FEMSolver
static void check(fsb_solver* s, int rc) {
  if (rc == FSB_OK) return;
  std::string msg = fsb_last_error(s);
  if (rc == FSB_ERR_INVALID) throw std::invalid_argument(msg);
  // upstream: FatalError -> backtrace + exit(1) (core/include/error.h:52-55)
  std::cerr << "FEMSolver (B200) fatal error: " << msg << std::endl;
  exit(1);
}

FEMSolver::FEMSolver(std::string fname, bool isTetMesh, bool verbose)
    : filename_(fname), tetMesh_(NULL), triMesh_(NULL), verbose_(verbose), device_(0),
      maxLevels_(100), topSize_(256), aggregatorType_(0), randMisParameters_(90102), partitionMaxSize_(512), proOmega_(0.67),
      preInnerIters_(5), postInnerIters_(5), postRelaxes_(1), smootherWeight_(1.0), dsType_(0),
      solverType_(0), maxIters_(100), tolerance_(1e-6),
      cycleIters_(1), cycleType_(0), convergeType_(0), blockSize_(256),
      seed_(0), refLevel0NoPerm_(0), iterations_(0), relres_(-1), impl_(NULL), A_from_file_(false) {
  if (fsb_create(&impl_, device_) != FSB_OK) {
    std::cerr << "FEMSolver (B200): " << fsb_last_error(NULL) << std::endl;
    exit(1);
  }
  if (isTetMesh) {
    this->tetMesh_ = TetMesh::read((this->filename_ + ".node").c_str(), (this->filename_ + ".ele").c_str(), verbose);
  } else {
    TriMesh::verbose = verbose;
    this->triMesh_ = TriMesh::read(this->filename_.c_str());
  }
  this->getMatrixFromMesh();
}

FEMSolver::~FEMSolver() {
  if (this->tetMesh_ != NULL) delete this->tetMesh_;
  if (this->triMesh_ != NULL) delete this->triMesh_;
  fsb_destroy(impl_);
}

void FEMSolver::pushParams() {
  struct { const char* n; double v; } p[] = {
    {"verbose", (double)verbose_}, {"maxLevels", (double)maxLevels_}, {"maxIters", (double)maxIters_},
    {"preInnerIters", (double)preInnerIters_}, {"postInnerIters", (double)postInnerIters_}, {"postRelaxes", (double)postRelaxes_},
    {"cycleIters", (double)cycleIters_}, {"dsType", (double)dsType_}, {"topSize", (double)topSize_},
    {"randMisParameters", (double)randMisParameters_}, {"partitionMaxSize", (double)partitionMaxSize_},
    {"aggregatorType", (double)aggregatorType_}, {"convergeType", (double)convergeType_}, {"tolerance", tolerance_},
    {"cycleType", (double)cycleType_}, {"solverType", (double)solverType_}, {"smootherWeight", smootherWeight_},
    {"proOmega", proOmega_}, {"blockSize", (double)blockSize_}, {"seed", (double)seed_}, {"refLevel0NoPerm", (double)refLevel0NoPerm_}};
  for (auto& q : p) check(impl_, fsb_set_param(impl_, q.n, q.v));
}

// device CSR (fp64) -> the public float ELL mirror: slot 0 = diagonal, then ascending neighbours,
// invalid_index padding (cutil.cu:197-226); also fills mesh->neighbors (sorted, as after tetmesh2ell).
void FEMSolver::pullMatrix() {
  int n = fsb_matrix_rows(impl_);
  long long nnz = fsb_matrix_nnz(impl_);
  std::vector<int> ptr(n + 1), col(nnz);
  std::vector<double> val(nnz);
  check(impl_, fsb_get_matrix_csr(impl_, ptr.data(), col.data(), val.data()));
  int maxrow = 0;
  for (int i = 0; i < n; i++) maxrow = std::max(maxrow, ptr[i + 1] - ptr[i]);
  A_h_.resize(n, n, nnz, maxrow, 32);
  std::vector<std::vector<int> >& nb = tetMesh_ ? tetMesh_->neighbors : triMesh_->neighbors;
  nb.assign(n, std::vector<int>());
  for (int i = 0; i < n; i++) {
    int slot = 1;
    for (int e = ptr[i]; e < ptr[i + 1]; e++) {
      if (col[e] == i) { A_h_.column_indices(i, 0) = i; A_h_.values(i, 0) = (float)val[e]; }
      else { A_h_.column_indices(i, slot) = col[e]; A_h_.values(i, slot) = (float)val[e]; nb[i].push_back(col[e]); slot++; }
    }
  }
}

void FEMSolver::getMatrixFromMesh() {
  if (this->triMesh_ == NULL && this->tetMesh_ == NULL) exit(0);  // FEMSolver.cu:142-143
  pushParams();
  if (tetMesh_) {
    size_t nv = tetMesh_->vertices.size(), ne = tetMesh_->tets.size();
    std::vector<double> xyz(3 * nv);
    std::vector<int> el(4 * ne);
    for (size_t i = 0; i < nv; i++) for (int j = 0; j < 3; j++) xyz[3 * i + j] = tetMesh_->vertices[i][j];
    for (size_t i = 0; i < ne; i++) for (int j = 0; j < 4; j++) el[4 * i + j] = tetMesh_->tets[i][j];
    check(impl_, fsb_set_tet_mesh(impl_, (int)nv, xyz.data(), (int)ne, el.data(), tetMesh_->matlabels.empty() ? NULL : tetMesh_->matlabels.data()));
  } else {
    size_t nv = triMesh_->vertices.size(), ne = triMesh_->faces.size();
    std::vector<double> xyz(3 * nv);
    std::vector<int> el(3 * ne);
    for (size_t i = 0; i < nv; i++) for (int j = 0; j < 3; j++) xyz[3 * i + j] = triMesh_->vertices[i][j];
    for (size_t i = 0; i < ne; i++) for (int j = 0; j < 3; j++) el[3 * i + j] = triMesh_->faces[i][j];
    check(impl_, fsb_set_tri_mesh(impl_, (int)nv, xyz.data(), (int)ne, el.data()));
  }
  check(impl_, fsb_assemble(impl_));
  pullMatrix();
  A_from_file_ = false;
}

size_t FEMSolver::getMatrixRows() { return this->A_h_.num_rows; }

void FEMSolver::checkMatrixForValidContents(Matrix_ell_h* A_h) {
  if (A_h->num_rows == 0) {
    if (this->verbose_) printf("Error no matrix specified\n");
    throw std::invalid_argument("Error no matrix specified");
  }
}

void FEMSolver::solveFEM(Vector_h_CG* x_h, Vector_h_CG* b_h) {
  this->checkMatrixForValidContents(&this->A_h_);
  pushParams();
  if (A_from_file_) {
    // the reference solves whatever is in A_h_ (float values, FEMSolver.cu:63): ELL -> CSR, ascending columns
    size_t n = A_h_.num_rows, K = A_h_.column_indices.num_cols;
    std::vector<int> ptr(n + 1, 0), col;
    std::vector<double> val;
    std::vector<std::pair<int, double> > row;
    for (size_t i = 0; i < n; i++) {
      row.clear();
      for (size_t j = 0; j < K; j++) {
        int c = A_h_.column_indices(i, j);
        if (c != Matrix_ell_h::invalid_index) row.push_back(std::make_pair(c, (double)A_h_.values(i, j)));
      }
      std::sort(row.begin(), row.end());
      for (auto& e : row) { col.push_back(e.first); val.push_back(e.second); }
      ptr[i + 1] = (int)col.size();
    }
    check(impl_, fsb_set_matrix_csr(impl_, (int)n, (long long)col.size(), ptr.data(), col.data(), val.data()));
  }
  int iters = 0; double relres = -1;
  check(impl_, fsb_solve_fem(impl_, b_h->data(), x_h->data(), &iters, &relres));
  iterations_ = iters; relres_ = relres;
  if (this->verbose_) printf("Computing time : %.10lf ms\n", fsb_time_ms(impl_, "setup") + fsb_time_ms(impl_, "solve"));
}

// ------------------------------------------------------------------------------------------- MATLAB v5 / VTK
namespace {
struct Reader {
  std::ifstream in;
  template <typename T> T get() { T v = T(); in.read((char*)&v, sizeof(T)); return v; }
  void skip(size_t n) { in.seekg((std::streamoff)n, std::ios::cur); }
};
// array-name element: small-data form or long form (FEMSolver.cu:232-256)
int skip_name(Reader& r, bool allow_long) {
  uint16_t type = r.get<uint16_t>();
  uint32_t len = r.get<uint16_t>();
  int align = 4;
  if (len == 0 && allow_long) { len = r.get<uint32_t>(); align = 8; }
  if (type != 1 && type != 2) {
    std::cerr << "WARNING: Invalid variable type (" << type << ") for array name characters (Must be 8-bit)." << std::endl;
    return -1;
  }
  if (len % align) len += align - len % align;
  r.skip(len);
  return 0;
}
}  // namespace

int FEMSolver::readMatlabSparseMatrix(const std::string& filename) {
  Reader r;
  r.in.open(filename.c_str(), std::ios::binary);
  if (!r.in.is_open()) { std::cerr << "could not open file: " << filename << std::endl; return 1; }
  r.skip(128);
  int32_t type = r.get<int32_t>();
  if (type == 15) { std::cerr << "Compression not supported. Save matlab data with '-v6' option." << std::endl; return 1; }
  if (type != 14) { std::cerr << filename << " is not a matlab matrix." << std::endl; return 1; }
  r.get<uint32_t>();
  type = r.get<int32_t>();
  if (type != 6 && type != 5) { std::cerr << "Invalid type for sparse matrix. Must be 32bit." << std::endl; return 1; }
  r.get<int32_t>();
  uint32_t mclass = r.get<uint32_t>() & 0xFF;
  if (mclass != 5) { std::cerr << "This is not a sparse matrix file." << std::endl; return 1; }
  r.get<uint32_t>();
  type = r.get<int32_t>();
  int32_t bytes = r.get<int32_t>();
  if ((type != 6 && type != 5) || bytes != 8) {
    std::cerr << "Matrix of wrong dimension type or # of dimensions." << std::endl;
    std::cerr << "Matrix must be 2 dimensions and of 32bit type." << std::endl;
    return 1;
  }
  int32_t x_dim = r.get<int32_t>(), y_dim = r.get<int32_t>();
  if (skip_name(r, true) != 0) return -1;
  auto read_ints = [&](std::vector<int32_t>& v, const char* what) -> bool {
    int32_t t = r.get<int32_t>();
    if (t != 6 && t != 5) { std::cerr << "Invalid " << what << " for sparse matrix. Must be 32bit." << std::endl; return false; }
    int32_t nb = r.get<int32_t>();
    v.assign(nb / 4, 0);
    r.in.read((char*)v.data(), nb);
    r.skip(nb % 8);
    return true;
  };
  std::vector<int32_t> row_vals, col_vals;
  if (!read_ints(row_vals, "type row index") || !read_ints(col_vals, "column index type")) return 1;
  type = r.get<int32_t>();
  if (type != 9) { std::cerr << "Invalid value for sparse matrix. Must be double float." << std::endl; return 1; }
  bytes = r.get<int32_t>();
  std::vector<double> vals(bytes / 8, 0);
  r.in.read((char*)vals.data(), bytes);
  // merge: every slot of the mesh pattern holds 1e-12f, the file entries are added on top (FEMSolver.cu:341-354)
  size_t n = (size_t)x_dim;
  if (n != A_h_.num_rows) { std::cerr << "matrix size does not match the mesh" << std::endl; return 1; }
  std::vector<std::vector<std::pair<int, float> > > rows(n);
  const std::vector<std::vector<int> >& nb = tetMesh_ ? tetMesh_->neighbors : triMesh_->neighbors;
  for (size_t i = 0; i < n; i++) {
    rows[i].push_back(std::make_pair((int)i, 1e-12f));
    for (int c : nb[i]) rows[i].push_back(std::make_pair(c, 1e-12f));
  }
  for (int32_t j = 0; j < y_dim; j++)
    for (int32_t q = col_vals[j]; q < col_vals[j + 1]; q++) {
      std::vector<std::pair<int, float> >& rw = rows[row_vals[q]];
      float v = static_cast<float>(vals[q]);
      bool found = false;
      for (auto& e : rw) if (e.first == j) { e.second = e.second + v; found = true; break; }
      if (!found) rw.push_back(std::make_pair((int)j, v));
    }
  size_t maxrow = 0, nnz = 0;
  for (auto& rw : rows) { std::sort(rw.begin(), rw.end()); maxrow = std::max(maxrow, rw.size()); nnz += rw.size(); }
  A_h_.resize(n, (size_t)y_dim, nnz, maxrow, 32);
  for (size_t i = 0; i < n; i++)
    for (size_t j = 0; j < rows[i].size(); j++) { A_h_.column_indices(i, j) = rows[i][j].first; A_h_.values(i, j) = rows[i][j].second; }
  A_from_file_ = true;
  return 0;
}

int FEMSolver::readMatlabArray(const std::string& filename, Vector_h_CG* rhs) {
  Reader r;
  r.in.open(filename.c_str(), std::ios::in | std::ios::binary);
  if (!r.in.is_open()) { std::cerr << "could not open file: " << filename << std::endl; return -1; }
  r.skip(128);
  int32_t type = r.get<int32_t>();
  if (type == 15) { std::cerr << "Compression not supported. Save matlab data with '-v6' option." << std::endl; return -1; }
  if (type != 14) { std::cerr << filename << " is not a matlab matrix." << std::endl; return -1; }
  r.get<uint32_t>();
  type = r.get<int32_t>();
  if (type != 6) { std::cerr << "Invalid type for normal matrix. Must be double precision." << std::endl; return -1; }
  r.get<int32_t>();
  uint32_t mclass = r.get<uint32_t>() & 0xFF;
  if (mclass == 5) { std::cerr << "This import routine is not for a sparse matrix file." << std::endl; return -1; }
  r.get<uint32_t>();
  type = r.get<int32_t>();
  int32_t bytes = r.get<int32_t>();
  if ((type != 6 && type != 5) || bytes != 8) {
    std::cerr << "Matrix of wrong dimension type or # of dimensions." << std::endl;
    return -1;
  }
  r.get<int32_t>(); r.get<int32_t>();
  if (skip_name(r, false) != 0) return -1;
  type = r.get<int32_t>();
  if (type != 9) { std::cerr << "Matrix data type must be miDOUBLE (type is " << type << ")." << std::endl; return -1; }
  uint32_t len = r.get<uint32_t>();
  std::vector<double> v(len / 8, 0);
  r.in.read((char*)v.data(), len);
  rhs->clear();
  for (size_t j = 0; j < v.size(); j++) rhs->push_back(v[j]);
  return 0;
}

int FEMSolver::writeMatlabArray(const std::string& filename, const Vector_h_CG& array) {
  std::ofstream file(filename.c_str(), std::ios::out | std::ios::binary);
  if (!file.is_open()) return 1;
  std::string desc = "MATLAB 5.0 MAT-file, Platform: GLNXA64, Created by SCI-Solver_FEM.";
  desc.resize(116, ' ');
  file.write(desc.c_str(), desc.length());
  char zeros[8] = {0};
  file.write(zeros, 8);
  int16_t version = 0x0100;
  file.write((char*)&version, 2);
  file.write("IM", 2);
  int32_t n = (int32_t)array.size();
  int32_t hdr[] = {14, 48 + n * 8, 6, 8, 6, 0, 5, 8, n, 1};
  file.write((char*)hdr, sizeof hdr);
  int16_t nameTag[] = {1, 3};
  file.write((char*)nameTag, 4);
  file.write("x_h\0", 4);
  int32_t dat[] = {9, n * 8};
  file.write((char*)dat, sizeof dat);
  for (size_t i = 0; i < array.size(); i++) { double v = array[i]; file.write((char*)&v, 8); }
  return 0;
}

void FEMSolver::writeVTK(std::vector<double> values, std::string fname) {
  FILE* f = fopen((fname + ".vtk").c_str(), "w+");
  if (!f) return;
  fprintf(f, "# vtk DataFile Version 3.0\nvtk output\nASCII\nDATASET UNSTRUCTURED_GRID\n");
  if (tetMesh_) {
    int nv = (int)tetMesh_->vertices.size(), nt = (int)tetMesh_->tets.size();
    fprintf(f, "POINTS %d float\n", nv);
    for (int i = 0; i < nv; i++) fprintf(f, "%.12f %.12f %.12f\n", tetMesh_->vertices[i][0], tetMesh_->vertices[i][1], tetMesh_->vertices[i][2]);
    fprintf(f, "CELLS %d %d\n", nt, nt * 5);
    for (int i = 0; i < nt; i++) fprintf(f, "4 %d %d %d %d\n", tetMesh_->tets[i][0], tetMesh_->tets[i][1], tetMesh_->tets[i][2], tetMesh_->tets[i][3]);
    fprintf(f, "CELL_TYPES %d\n", nt);
    for (int i = 0; i < nt; i++) fprintf(f, "10\n");
    fprintf(f, "POINT_DATA %d\nSCALARS traveltime float 1\nLOOKUP_TABLE default\n", nv);
    for (size_t i = 0; i < values.size(); i++) fprintf(f, "%.12f\n ", values[i]);
  } else if (triMesh_) {
    int nv = (int)triMesh_->vertices.size(), nt = (int)triMesh_->faces.size();
    fprintf(f, "POINTS %d float\n", nv);
    for (int i = 0; i < nv; i++) fprintf(f, "%.12f %.12f %.12f\n", triMesh_->vertices[i][0], triMesh_->vertices[i][1], triMesh_->vertices[i][2]);
    fprintf(f, "CELLS %d %d\n", nt, nt * 4);
    for (int i = 0; i < nt; i++) fprintf(f, "3 %d %d %d\n", triMesh_->faces[i][0], triMesh_->faces[i][1], triMesh_->faces[i][2]);
    fprintf(f, "CELL_TYPES %d\n", nt);
    for (int i = 0; i < nt; i++) fprintf(f, "5\n");
    fprintf(f, "POINT_DATA %d\nSCALARS traveltime float 1\nLOOKUP_TABLE default\n", nv);
    for (int i = 0; i < nv && i < (int)values.size(); i++) fprintf(f, "%.12f\n", static_cast<float>(values[i]));
  }
  fclose(f);
}
