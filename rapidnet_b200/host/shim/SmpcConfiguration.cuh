// SmpcConfiguration.cuh (shim) -- the reference's header name (/root/reference/src/SmpcConfiguration.cuh) for callers compiled against rapidnet-b200:
// class SmpcConfiguration of rapidnet_b200/host/rapidnet_host.hpp in the global namespace, where the reference declares it.
#pragma once
#include "Configuration.h"
using rapidnet::SmpcConfiguration;
// keys of the JSON document this class loads (the reference's macros, /root/reference/src/SmpcConfiguration.cuh:24-47: callers and the
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
#ifndef VARNAME_NV
#define VARNAME_NV "nv"
#endif
#ifndef VARNAME_N
#define VARNAME_N "N"
#endif
#ifndef VARNAME_L
#define VARNAME_L "matL"
#endif
#ifndef VARNAME_LHAT
#define VARNAME_LHAT "matLhat"
#endif
#ifndef VARNAME_COSTW
#define VARNAME_COSTW "costW"
#endif
#ifndef VARNAME_PENALITY_X
#define VARNAME_PENALITY_X "penaltyStateX"
#endif
#ifndef VARNAME_PENALITY_XS
#define VARNAME_PENALITY_XS "penaltySafetyX"
#endif
#ifndef VARNAME_DIAG_PRCND
#define VARNAME_DIAG_PRCND "matDiagPrecnd"
#endif
#ifndef VARNAME_CURRENT_X
#define VARNAME_CURRENT_X "currentX"
#endif
#ifndef VARNAME_PREV_UHAT
#define VARNAME_PREV_UHAT "prevUhat"
#endif
#ifndef VARNAME_PREV_U
#define VARNAME_PREV_U "prevU"
#endif
#ifndef VARNAME_PREV_V
#define VARNAME_PREV_V "prevV"
#endif
#ifndef VARNAME_PREV_DEMAND
#define VARNAME_PREV_DEMAND "prevDemand"
#endif
#ifndef VARNAME_STEP_SIZE
#define VARNAME_STEP_SIZE "stepSize"
#endif
#ifndef VARNAME_MAX_ITER
#define VARNAME_MAX_ITER "maxIterations"
#endif
#ifndef VARNAME_LBFGS_BUFFER_SIZE
#define VARNAME_LBFGS_BUFFER_SIZE "lbfgsBufferSize"
#endif
#ifndef PATH_NETWORK_FILE
#define PATH_NETWORK_FILE "pathToNetwork"
#endif
#ifndef PATH_SCENARIO_TREE_FILE
#define PATH_SCENARIO_TREE_FILE "pathToScenarioTree"
#endif
#ifndef PATH_FORECASTER_FILE
#define PATH_FORECASTER_FILE "pathToForecaster"
#endif
