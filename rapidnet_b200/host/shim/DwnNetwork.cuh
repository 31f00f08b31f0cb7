// DwnNetwork.cuh (shim) -- the reference's header name (/root/reference/src/DwnNetwork.cuh) for callers compiled against rapidnet-b200:
// class DwnNetwork of rapidnet_b200/host/rapidnet_host.hpp in the global namespace, where the reference declares it.
#pragma once
#include "Configuration.h"
using rapidnet::DwnNetwork;
// keys of the JSON document this class loads (the reference's macros, /root/reference/src/DwnNetwork.cuh:23-37: callers and the
// reference's tests spell the keys through them); repeated definitions across the loader headers are identical, as there
#ifndef VARNAME_NX
#define VARNAME_NX "nx"
#endif
#ifndef VARNAME_NU
#define VARNAME_NU "nu"
#endif
#ifndef VARNAME_ND
#define VARNAME_ND "nd"
#endif
#ifndef VARNAME_NE
#define VARNAME_NE "ne"
#endif
#ifndef VARNAME_A
#define VARNAME_A "matA"
#endif
#ifndef VARNAME_B
#define VARNAME_B "matB"
#endif
#ifndef VARNAME_GD
#define VARNAME_GD "matGd"
#endif
#ifndef VARNAME_E
#define VARNAME_E "matE"
#endif
#ifndef VARNAME_ED
#define VARNAME_ED "matEd"
#endif
#ifndef VARNAME_XMIN
#define VARNAME_XMIN "vecXmin"
#endif
#ifndef VARNAME_XMAX
#define VARNAME_XMAX "vecXmax"
#endif
#ifndef VARNAME_XSAFE
#define VARNAME_XSAFE "vecXsafe"
#endif
#ifndef VARNAME_UMIN
#define VARNAME_UMIN "vecUmin"
#endif
#ifndef VARNAME_UMAX
#define VARNAME_UMAX "vecUmax"
#endif
#ifndef VARNAME_ALPHA1
#define VARNAME_ALPHA1 "costAlpha1"
#endif
