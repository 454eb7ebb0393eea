#pragma once
#include <x86intrin.h>
#include <mmintrin.h>
