#pragma once
#include "windows.h"
