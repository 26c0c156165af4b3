// types.h — CUSP-free, source-compatible stand-ins for the types that cross the FEMSolver boundary.
// Reference: src/core/include/types.h:28 (Vector_h_CG = cusp::array1d<double, host_memory>) and :51
// (Matrix_ell_h = cusp::ell_matrix<int, float, host_memory>).  Callers use the (n, value)
// constructor, operator[], size(); the library additionally uses clear() / push_back()
// (src/FEMSolver.cu:444-447).
#ifndef __TYPES_H__
#define __TYPES_H__
#include <cstddef>
#include <vector>

typedef double CGType;
typedef float AMGType;
typedef double AssembleType;

template <typename T>
class fsb_array1d : public std::vector<T> {
 public:
  fsb_array1d() {}
  explicit fsb_array1d(size_t n) : std::vector<T>(n) {}
  fsb_array1d(size_t n, const T& v) : std::vector<T>(n, v) {}
  template <typename It>
  fsb_array1d(It a, It b) : std::vector<T>(a, b) {}
};
typedef fsb_array1d<double> Vector_h_CG;
typedef fsb_array1d<float> Vector_h;
typedef fsb_array1d<int> IdxVector_h;

// ELL matrix with cusp's accessors: column_indices(i, j), values(i, j), invalid_index padding.
template <typename IndexType, typename ValueType>
class fsb_ell_matrix {
 public:
  template <typename T>
  struct array2d {
    size_t num_rows = 0, num_cols = 0, pitch = 0;
    std::vector<T> values;  // column-major with pitch, as cusp::array2d<..., column_major>
    void resize(size_t r, size_t c, size_t p) { num_rows = r; num_cols = c; pitch = p; values.assign(p * c, T()); }
    T& operator()(size_t i, size_t j) { return values[j * pitch + i]; }
    const T& operator()(size_t i, size_t j) const { return values[j * pitch + i]; }
  };
  typedef IndexType index_type;
  typedef ValueType value_type;
  static const IndexType invalid_index = static_cast<IndexType>(-1);
  size_t num_rows = 0, num_cols = 0, num_entries = 0;
  array2d<IndexType> column_indices;
  array2d<ValueType> values;
  fsb_ell_matrix() {}
  fsb_ell_matrix(size_t r, size_t c, size_t nnz, size_t per_row, size_t alignment = 32) { resize(r, c, nnz, per_row, alignment); }
  void resize(size_t r, size_t c, size_t nnz, size_t per_row, size_t alignment = 32) {
    num_rows = r; num_cols = c; num_entries = nnz;
    size_t pitch = alignment * ((r + alignment - 1) / alignment);
    column_indices.resize(r, per_row, pitch);
    values.resize(r, per_row, pitch);
    for (size_t k = 0; k < column_indices.values.size(); k++) column_indices.values[k] = invalid_index;
  }
};
typedef fsb_ell_matrix<int, float> Matrix_ell_h;
typedef fsb_ell_matrix<int, double> Matrix_ell_h_CG;
#endif
