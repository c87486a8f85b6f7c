/* Stand-ins for the three esp-dsp (^1.7.0, not vendored) calls mfcc.c makes (mfcc.c:259-261,346-348).
 * They realise the INTENDED math: after the three calls fft_input[2k], fft_input[2k+1] hold
 * Re/Im of the 512-point DFT of the (zero-imaginary) frame.  PARITY UNPINNED for this stage. */
#pragma once
int dsps_fft2r_fc32_ansi(float* data, int n);
int dsps_bit_rev_fc32(float* data, int n);
int dsps_cplx2reC_fc32(float* data, int n);
