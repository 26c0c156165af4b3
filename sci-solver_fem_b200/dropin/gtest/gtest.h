// Minimal googletest-compatible shim (TEST, ASSERT_TRUE/EQ/LT, EXPECT_TRUE, main) so the reference's
// src/test/*.cc build unchanged: googletest is fetched from the network by the upstream CMake
// (src/test/CMakeLists.txt:9-20) and is not available offline.
#ifndef FSB_GTEST_SHIM_H
#define FSB_GTEST_SHIM_H
#include <cmath>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>
#include <cstdlib>
// TEST_DATA_DIR is a compile definition upstream (src/test/CMakeLists.txt:33); here it expands to this
// call so the same binaries can be pointed at materialised fixtures with FSB_TEST_DATA_DIR.
inline const char* fsb_test_data_dir() {
  const char* e = std::getenv("FSB_TEST_DATA_DIR");
  return e ? e : "/root/reference/src/test/test_data";
}
namespace testing {
struct TestInfo { const char* suite; const char* name; void (*fn)(bool&); };
inline std::vector<TestInfo>& registry() { static std::vector<TestInfo> r; return r; }
struct Registrar { Registrar(const char* s, const char* n, void (*f)(bool&)) { registry().push_back({s, n, f}); } };
inline void InitGoogleTest(int*, char**) {}
}  // namespace testing
#define TEST(suite, name)                                                        \
  static void suite##_##name##_body(bool& gtest_failed);                          \
  static ::testing::Registrar suite##_##name##_reg(#suite, #name, suite##_##name##_body); \
  static void suite##_##name##_body(bool& gtest_failed)
#define FSB_CHECK_(cond, fatal)                                                               \
  do { if (!(cond)) { std::cerr << __FILE__ << ":" << __LINE__ << ": Failure\n  " #cond << std::endl; gtest_failed = true; if (fatal) return; } } while (0)
#define ASSERT_TRUE(c) FSB_CHECK_((c), true)
#define EXPECT_TRUE(c) FSB_CHECK_((c), false)
#define ASSERT_EQ(a, b) FSB_CHECK_((a) == (b), true)
#define ASSERT_LT(a, b) FSB_CHECK_((a) < (b), true)
inline int RUN_ALL_TESTS() {
  int failed = 0;
  for (auto& t : ::testing::registry()) {
    std::cout << "[ RUN      ] " << t.suite << "." << t.name << std::endl;
    bool f = false;
    t.fn(f);
    std::cout << (f ? "[  FAILED  ] " : "[       OK ] ") << t.suite << "." << t.name << std::endl;
    failed += f;
  }
  return failed ? 1 : 0;
}
#ifndef FSB_GTEST_NO_MAIN
int main(int argc, char** argv) { ::testing::InitGoogleTest(&argc, argv); return RUN_ALL_TESTS(); }
#endif
#endif
