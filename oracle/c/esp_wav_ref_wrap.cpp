// TEST INFRASTRUCTURE ONLY.  extern "C" window onto the reference's own wav::WavHeader
// (main/esp_wav/esp_wav.cpp + esp_wav.hpp, compiled from /root/reference by oracle/c/Makefile into
// oracle/_ref/libesp_wav_ref.so).  Used to pin oracle/wav.py and to generate tests/golden/wav_cases.npz.
#include <cstdint>
#include <cstring>

#include "esp_wav.hpp"

extern "C" int esp_wav_ref_parse(const char* path, uint32_t* fields /* [10] */) {
    wav::WavHeader h{std::string(path)};
    fields[0] = h.getNumChannels();
    fields[1] = h.getSampleRate();
    fields[2] = h.getBitsPerSample();
    fields[3] = h.getDataLength();
    fields[4] = h.getByteRate();
    fields[5] = h.getBlockAlign();
    fields[6] = h.getRawDataPosition();
    fields[7] = h.isValid() ? 1u : 0u;
    fields[8] = (uint32_t)h.getFileStatus();
    fields[9] = 0;
    return 0;
}

// writer side: WavHeader(path, channels, sr, bps) + write_info_to_file + write_data_to_file + finalize_wav_file
extern "C" int esp_wav_ref_write(const char* path, const int16_t* pcm, uint32_t n, uint16_t channels, uint32_t sr) {
    wav::WavHeader h(std::string(path), channels, sr, 16, 0);
    if (!h.write_info_to_file()) return -1;
    if (n && !h.write_data_to_file(pcm, n)) return -2;
    return h.finalize_wav_file() ? 0 : -3;
}
