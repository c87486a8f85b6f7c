#pragma once
#include <stdio.h>
/* TEST INFRASTRUCTURE ONLY.  The reference's log calls: errors go to stderr; the last info / error line is also
 * kept in a buffer that the golden generators read back (tests/golden/make_golden_range.py), so that what the
 * reference LOGS (analyze_mfcc_range, mfcc.c:530-553) can be pinned as well. */
__attribute__((weak)) char shim_last_info[512];
__attribute__((weak)) char shim_last_error[512];
#define ESP_LOGE(tag, ...) do { fprintf(stderr, "E %s: ", tag); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); \
                                snprintf(shim_last_error, sizeof shim_last_error, __VA_ARGS__); } while (0)
#define ESP_LOGI(tag, ...) do { (void)(tag); snprintf(shim_last_info, sizeof shim_last_info, __VA_ARGS__); } while (0)
#define ESP_LOGW(tag, ...) do { (void)(tag); } while (0)
