/* empty: interp.cpp:1 / pipeline.cpp:2 include ATL but use nothing from it */
#pragma once
