// Forecaster.cuh (shim) -- the reference's header name (/root/reference/src/Forecaster.cuh) for callers compiled against rapidnet-b200:
// class Forecaster of rapidnet_b200/host/rapidnet_host.hpp in the global namespace, where the reference declares it.
#pragma once
#include "Configuration.h"
using rapidnet::Forecaster;
// keys of the JSON document this class loads (the reference's macros, /root/reference/src/Forecaster.cuh:23-30: callers and the
// reference's tests spell the keys through them); repeated definitions across the loader headers are identical, as there
#ifndef VARNAME_N
#define VARNAME_N "N"
#endif
#ifndef VARNAME_SIM_HORIZON
#define VARNAME_SIM_HORIZON "simHorizon"
#endif
#ifndef VARNAME_DIM_DEMAND
#define VARNAME_DIM_DEMAND "dimDemand"
#endif
#ifndef VARNAME_DIM_PRICES
#define VARNAME_DIM_PRICES "dimPrices"
#endif
#ifndef VARNAME_DHAT
#define VARNAME_DHAT "dHat"
#endif
#ifndef VARNAME_ALPHAHAT
#define VARNAME_ALPHAHAT "alphaHat"
#endif
#ifndef VARNAME_DEMAND_SIM
#define VARNAME_DEMAND_SIM "timeIdDemand4876"
#endif
#ifndef VARNAME_PRICE_SIM
#define VARNAME_PRICE_SIM "timeIdPrice4876"
#endif
