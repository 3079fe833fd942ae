/* include/paut.h consumed from plain C99: the header must compile without C++, every entry point must resolve, and
 * the host-only calls must work without a GPU.  Built and run by tests/test_abi.py::test_header_is_valid_c. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "paut.h"

#define LOAD(name)                                                  \
  do {                                                              \
    *(void**)(&name##_p) = dlsym(lib, #name);                       \
    if (!name##_p) { fprintf(stderr, "missing %s\n", #name); return 2; } \
  } while (0)

int main(int argc, char** argv) {
  int (*paut_abi_version_p)(void);
  int (*paut_ctx_create_p)(int, void*, paut_ctx**);
  const char* (*paut_last_error_p)(const paut_ctx*);
  int (*paut_window_table_host_p)(int, int64_t, int64_t, int32_t*, int);
  int (*paut_json_load_host_p)(const char*, paut_json_volume**);
  int (*paut_json_num_beams_p)(const paut_json_volume*);
  int (*paut_json_beam_info_p)(const paut_json_volume*, int, const char**, int64_t*, int64_t*);
  void (*paut_json_free_p)(paut_json_volume*);
  void* lib;
  int32_t pairs[16];
  int n, has_gpu;
  paut_ctx* ctx = NULL;
  paut_json_volume* vol = NULL;
  paut_detection det;
  paut_metrics met;
  if (argc < 3) return 1;
  lib = dlopen(argv[1], RTLD_NOW);
  if (!lib) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  LOAD(paut_abi_version); LOAD(paut_ctx_create); LOAD(paut_last_error); LOAD(paut_window_table_host);
  LOAD(paut_json_load_host); LOAD(paut_json_num_beams); LOAD(paut_json_beam_info); LOAD(paut_json_free);
  if (paut_abi_version_p() != PAUT_ABI_VERSION) return 3;
  if (sizeof(det) != 48 || sizeof(met) != 48) return 4;
  n = paut_window_table_host_p(1, 120, 50, pairs, 8);          /* starts 0,17,34,51,68 + tail 70 */
  if (n != 6 || pairs[2] != 17 || pairs[10] != 70) return 5;
  n = paut_window_table_host_p(0, 120, 50, pairs, 8);          /* starts 0,50,70 */
  if (n != 3 || pairs[4] != 70) return 6;
  if (paut_json_load_host_p(argv[2], &vol) != PAUT_OK) return 7;
  {
    const char* key; int64_t scans, S;
    if (paut_json_num_beams_p(vol) != 3) return 8;
    if (paut_json_beam_info_p(vol, 0, &key, &scans, &S) != PAUT_OK || strcmp(key, "beam_0") || scans != 13 || S != 8) return 9;
  }
  paut_json_free_p(vol);
  has_gpu = paut_ctx_create_p(0, NULL, &ctx) == PAUT_OK;
  if (!has_gpu && !strstr(paut_last_error_p(NULL), "no CUDA device")) return 10;   /* no CPU fallback */
  printf("c abi ok (gpu %d)\n", has_gpu);
  return 0;
}
