#pragma once
#include <stddef.h>
#include <stdint.h>
