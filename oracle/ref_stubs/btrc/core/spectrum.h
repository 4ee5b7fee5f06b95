// ORACLE build shim (test infrastructure).  The reference's <btrc/core/spectrum.h> declares the host class Spectrum *and* its CUJ
// device proxy (CUJ_PROXY_CLASS_EX), which gcc cannot compile; factory/node/node.cpp only needs the host class to exist so that
// Node::parse<Spectrum> can return one.  This header, found first on the include path of the `ref` build, gives it the four
// floats and the one factory function node.cpp calls; node.cpp itself is compiled where it lies, unmodified.
#pragma once
#include <btrc/utils/math/math.h>
BTRC_BEGIN
class Spectrum
{
public:
    float r = 0, g = 0, b = 0, additional_data = 0;
    static Spectrum from_rgb(float r_, float g_, float b_) { Spectrum s; s.r = r_; s.g = g_; s.b = b_; return s; }
};
BTRC_END
