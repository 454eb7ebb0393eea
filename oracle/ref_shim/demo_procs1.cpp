// Factory for the shader classes of demo 1 (src/test/testproc.h), compiled against the patched copy in the scratch dir.
// TEST INFRASTRUCTURE (part of oracle/_ref/libps3d_ref.so).
#include "windows.h"
#include "pipeline.h"
#include "testproc.h"
#include "ps3d.h"

PuresoftProcessor* ps3d_demo1_make_processor(int kind, int functor)
{
	switch(functor)
	{
	case PS3D_FN_PLANET:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_Planet : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new IP_Planet : (PuresoftProcessor*)new FP_Earth;
	case PS3D_FN_SATELLITE:
		return kind == PS3D_PROC_FRAGMENT ? (PuresoftProcessor*)new FP_Satellite : NULL;
	case PS3D_FN_CLOUD:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_Cloud : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new IP_Cloud : (PuresoftProcessor*)new FP_Cloud;
	case PS3D_FN_CLOUDSHADOW:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_CloudShadow : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new IP_CloudShadow : (PuresoftProcessor*)new FP_CloudShadow;
	}
	return NULL;
}
