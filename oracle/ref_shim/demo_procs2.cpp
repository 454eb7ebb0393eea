// Factory for the shader classes of demo 2 (src/test2/testproc.h), compiled against the patched copy in the scratch dir.
// TEST INFRASTRUCTURE (part of oracle/_ref/libps3d_ref.so).
#include "windows.h"
#include "pipeline.h"
#include "testproc.h"
#include "testpost.h"
#include "ps3d.h"

PuresoftProcessor* ps3d_demo2_make_processor(int kind, int functor)
{
	switch(functor)
	{
	case PS3D_FN_POSITIONONLY:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_PositionOnly : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new IP_Null : (PuresoftProcessor*)new FP_SingleColourNoLighting;
	case PS3D_FN_SINGLECOLOUR:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_SingleColour : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new IP_SingleColour : (PuresoftProcessor*)new FP_SingleColour;
	case PS3D_FN_DIFFUSEONLY:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_DiffuseOnly : kind == PS3D_PROC_INTERPOLATION ? (PuresoftProcessor*)new IP_DiffuseOnly : (PuresoftProcessor*)new FP_DiffuseOnly;
	case PS3D_FN_SHADOW2:
		return kind == PS3D_PROC_VERTEX ? (PuresoftProcessor*)new VP_Shadow : kind == PS3D_PROC_FRAGMENT ? (PuresoftProcessor*)new FP_Null : NULL;
	}
	return NULL;
}

PuresoftPostProcessor* ps3d_demo2_make_post_processor(int functor)
{
	return functor == PS3D_POST_DEPTHOFFIELD ? new PP_DepthofField : NULL;
}
