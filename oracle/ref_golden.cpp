// TEST INFRASTRUCTURE ONLY (oracle/_ref). Exposes the reference's own JFastFir
// golden vectors (JAERO/tests/jfastfir_data_{input,expected_output}.cpp, compiled
// verbatim where they lie) so tests/ and tools/ can read them through ctypes.
#include "qt_shim.h"
#include "DSP.h"
extern const QVector<cpx_type> input;
extern const QVector<cpx_type> expected_output;
extern double Fs;
extern double fb;
extern "C" {
long jref_golden_jfastfir_len(void) { return (long)input.size(); }
double jref_golden_jfastfir_Fs(void) { return Fs; }
double jref_golden_jfastfir_fb(void) { return fb; }
void jref_golden_jfastfir(double *in_ri, double *out_ri)
{
    for (int i = 0; i < input.size(); i++) {
        in_ri[2 * i] = input[i].real(); in_ri[2 * i + 1] = input[i].imag();
        out_ri[2 * i] = expected_output[i].real(); out_ri[2 * i + 1] = expected_output[i].imag();
    }
}
}
