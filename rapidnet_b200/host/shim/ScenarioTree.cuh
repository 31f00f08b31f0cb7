// ScenarioTree.cuh (shim) -- the reference's header name (/root/reference/src/ScenarioTree.cuh) for callers compiled against rapidnet-b200:
// class ScenarioTree of rapidnet_b200/host/rapidnet_host.hpp in the global namespace, where the reference declares it.
#pragma once
#include "Configuration.h"
using rapidnet::ScenarioTree;
// keys of the JSON document this class loads (the reference's macros, /root/reference/src/ScenarioTree.cuh:23-40: callers and the
// reference's tests spell the keys through them); repeated definitions across the loader headers are identical, as there
#ifndef VARNAME_N
#define VARNAME_N "N"
#endif
#ifndef VARNAME_K
#define VARNAME_K "K"
#endif
#ifndef VARNAME_NODES
#define VARNAME_NODES "nodes"
#endif
#ifndef VARNAME_NUM_NONLEAF
#define VARNAME_NUM_NONLEAF "nNonLeafNodes"
#endif
#ifndef VARNAME_NUM_CHILD_TOT
#define VARNAME_NUM_CHILD_TOT "nChildrenTot"
#endif
#ifndef VARNAME_STAGES
#define VARNAME_STAGES "stages"
#endif
#ifndef VARNAME_NODES_PER_STAGE
#define VARNAME_NODES_PER_STAGE "nodesPerStage"
#endif
#ifndef VARNAME_NODES_PER_STAGE_CUMUL
#define VARNAME_NODES_PER_STAGE_CUMUL "nodesPerStageCumul"
#endif
#ifndef VARNAME_LEAVES
#define VARNAME_LEAVES "leaves"
#endif
#ifndef VARNAME_CHILDREN
#define VARNAME_CHILDREN "children"
#endif
#ifndef VARNAME_ANCESTOR
#define VARNAME_ANCESTOR "ancestor"
#endif
#ifndef VARNAME_NUM_CHILDREN
#define VARNAME_NUM_CHILDREN "nChildren"
#endif
#ifndef VARNAME_NUM_CHILD_CUMUL
#define VARNAME_NUM_CHILD_CUMUL "nChildrenCumul"
#endif
#ifndef VARNAME_PROB_NODE
#define VARNAME_PROB_NODE "probNode"
#endif
#ifndef VARNAME_DIM_DEMAND
#define VARNAME_DIM_DEMAND "dimDemand"
#endif
#ifndef VARNAME_DIM_PRICE
#define VARNAME_DIM_PRICE "dimPrice"
#endif
#ifndef VARNAME_DEMAND_NODE
#define VARNAME_DEMAND_NODE "errorDemandNode"
#endif
#ifndef VARNAME_PRICE_NODE
#define VARNAME_PRICE_NODE "errorPriceNode"
#endif
