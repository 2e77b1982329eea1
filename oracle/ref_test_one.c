/*
 * oracle/ref_test_one.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Runs ONE of the reference's own test functions (reference test/test_qc_*.c, compiled unchanged
 * from where they lie under /root/reference/test by oracle/Makefile `reftests`) so that a harness
 * can record the pass / fail status of each of the 20 separately: the reference's own driver
 * (test/test_main.c:26-59) stops at the first failing assert, and the unmodified reference fails
 * two of its tests.
 *
 *   ref_tests_<lib> <test name | all | list> [seed]
 *
 * Exit status 0 = the test's asserts held; an assert failure aborts (SIGABRT) like in the
 * reference's driver.  The same object files are linked three ways: against libqcs.so (this
 * repository: QCS_GPU_CUDA), against the unmodified reference and against the corrected reference.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void test_qc_create_destroy(void);
void test_qc_getters(void);
void test_qc_h(void);
void test_qc_x(void);
void test_qc_y(void);
void test_qc_z(void);
void test_qc_cnot(void);
void test_qc_phase(void);
void test_qc_rotations(void);
void test_qc_barrier(void);
void test_qc_reset(void);
void test_qc_measure(void);
void test_qc_run(void);
void test_qc_run_shots(void);
void test_qc_state_access(void);
void test_qc_print(void);
void test_qc_grover_search(void);
void test_qc_qft(void);
void test_qc_bv(void);
void test_qc_optimize(void);

struct entry {
  const char *name;
  void (*fn)(void);
};

/* the order of the reference's driver (test/test_main.c:34-53) */
static const struct entry table[] = {
    {"test_qc_create_destroy", test_qc_create_destroy},
    {"test_qc_getters", test_qc_getters},
    {"test_qc_h", test_qc_h},
    {"test_qc_x", test_qc_x},
    {"test_qc_y", test_qc_y},
    {"test_qc_z", test_qc_z},
    {"test_qc_cnot", test_qc_cnot},
    {"test_qc_phase", test_qc_phase},
    {"test_qc_rotations", test_qc_rotations},
    {"test_qc_barrier", test_qc_barrier},
    {"test_qc_reset", test_qc_reset},
    {"test_qc_measure", test_qc_measure},
    {"test_qc_run", test_qc_run},
    {"test_qc_run_shots", test_qc_run_shots},
    {"test_qc_state_access", test_qc_state_access},
    {"test_qc_print", test_qc_print},
    {"test_qc_grover_search", test_qc_grover_search},
    {"test_qc_qft", test_qc_qft},
    {"test_qc_bv", test_qc_bv},
    {"test_qc_optimize", test_qc_optimize},
};

int main(int argc, char **argv) {
  const int n = (int)(sizeof(table) / sizeof(table[0]));
  int i, ran = 0;
  if (argc < 2) {
    fprintf(stderr, "usage: %s <test name | all | list> [seed]\n", argv[0]);
    return 2;
  }
  srand(argc > 2 ? (unsigned)atoi(argv[2]) : 12345u);
  for (i = 0; i < n; i++) {
    if (strcmp(argv[1], "list") == 0) {
      printf("%s\n", table[i].name);
      ran++;
    } else if (strcmp(argv[1], "all") == 0 || strcmp(argv[1], table[i].name) == 0) {
      table[i].fn();
      fflush(stdout);
      ran++;
    }
  }
  if (!ran) {
    fprintf(stderr, "unknown test %s\n", argv[1]);
    return 2;
  }
  return 0;
}
