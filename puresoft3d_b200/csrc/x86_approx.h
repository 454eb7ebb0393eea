// x86_approx.h — host-side measurement of rcpps / rsqrtss (see x86_approx.cpp)
#pragma once
#include <stdint.h>
#include <vector>

struct Ps3dHostApprox
{
	std::vector<uint32_t> rcp, rsqrt;
	int rcpBits, rsqrtBits;
};

// true: tables filled and verified against this CPU; false: use correctly rounded 1/x and 1/sqrt(x)
bool ps3d_measure_x86_approx(Ps3dHostApprox* out);
