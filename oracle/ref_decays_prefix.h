/* oracle/ref_decays_prefix.h -- TEST INFRASTRUCTURE.  Force-included (g++ -include) in front of the reference's
 * emissionfunction_resonance_decays.cpp, and of that translation unit only, when oracle/_ref/is3d_ref_decays is built.
 *
 * EmissionFunctionArray::do_resonance_decays prints "I need to change the linear interpolation's MTmax ..." and calls exit(-1) as
 * its first statements (emissionfunction_resonance_decays.cpp:126-129), so the stock reference never runs the routine below it.
 * The body is nevertheless the only specification of SURVEY 8f row N3.  This prefix lets it run WITHOUT editing the source: the
 * standard headers are included first (so the real, noreturn exit() is declared untouched), then `exit(code)` in the rest of the
 * translation unit is routed to a hook defined in oracle/ref_driver.cpp that swallows exactly the first call and really exits on any
 * later one (the routine's genuine error paths).  Vectors made this way are labelled "reference author flags the MTmax handling of
 * the interpolation as unfinished". */
#include <stdlib.h>
#include <cstdlib>
#ifdef __cplusplus
extern "C"
#endif
void is3d_ref_exit_hook(int code);
#define exit(code) is3d_ref_exit_hook(code)
