// Engine.cuh (shim) -- the reference's header name (/root/reference/src/Engine.cuh) for callers compiled against rapidnet-b200:
// class Engine of rapidnet_b200/host/rapidnet_host.hpp in the global namespace, where the reference declares it.
#pragma once
#include "Configuration.h"
#include "DwnNetwork.cuh"
#include "ScenarioTree.cuh"
#include "SmpcConfiguration.cuh"
#include "Forecaster.cuh"
#include "Utilities.cuh"
#include "cublas_v2.h"
using rapidnet::Engine;
