#pragma once
#include <stdio.h>
#define ESP_LOGE(tag, ...) do { fprintf(stderr, "E %s: ", tag); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ESP_LOGI(tag, ...) do { (void)(tag); } while (0)
#define ESP_LOGW(tag, ...) do { (void)(tag); } while (0)
