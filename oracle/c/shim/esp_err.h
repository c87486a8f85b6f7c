#pragma once
/* shim for building reference sources on the host -- TEST INFRASTRUCTURE ONLY */
typedef int esp_err_t;
#define ESP_OK 0
#define ESP_FAIL -1
