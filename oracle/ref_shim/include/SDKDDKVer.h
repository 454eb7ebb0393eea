/* empty: src/mcemath/targetver.h includes the Windows SDK version header */
