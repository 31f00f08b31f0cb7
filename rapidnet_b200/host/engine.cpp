// engine.cpp -- Engine and SmpcController of the reference as thin host code over the C ABI.
//   Engine          /root/reference/src/Engine.cu:126-163 (ctor), :671-774 (factorStep), :1147-1316 (per-solve terms)
//   SmpcController  /root/reference/src/SmpcController.cu:32-116 (ctors), :476-487, :535-864 (steps), :1500-1525 (APG),
//                   :1593-1667 (controller calls), :1679-1717 (moveForewardInTime), :1778-1859 (KPIs)
// Error behaviour as in the reference (_CUDA / _CUBLAS / _ASSERT, src/Configuration.h:38-81): print and exit.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <iostream>

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include "rapidnet_host.hpp"

namespace rapidnet {

// ---- Engine ------------------------------------------------------------------------------------------------------------
void Engine::check(rn_status rc, const char *what) {
    if (rc == RN_OK) return;
    std::cerr << "rapidnet_b200: " << what << " failed (" << rc << "): " << rn_last_error(h) << std::endl;
    std::exit(EXIT_FAILURE);
}

Engine::Engine(SmpcConfiguration *smpcConfig) {
    ptrMySmpcConfig = smpcConfig;
    ptrMyNetwork = new DwnNetwork(smpcConfig->getPathToNetwork());
    ptrMyScenarioTree = new ScenarioTree(smpcConfig->getPathToScenarioTree());
    const std::string algorithmName = smpcConfig->getOptimisationAlgorithm();
    std::cout << "Number of scenarios " << ptrMyScenarioTree->getNumScenarios() << std::endl;
    globalFbeFlag = algorithmName == "globalFbeAlgorithm";
    namaFlag = algorithmName == "namaAlgorithm";
    apgFlag = !globalFbeFlag && !namaFlag;
    std::cout << "algorithm based on SMPC " << algorithmName << std::endl;

    DwnNetwork *n = ptrMyNetwork;
    ScenarioTree *t = ptrMyScenarioTree;
    rn_dims d{};
    d.nx = n->getNumTanks(); d.nu = n->getNumControls(); d.nd = n->getNumDemands(); d.ne = n->getNumMixNodes();
    d.nv = smpcConfig->getNV(); d.N = t->getPredHorizon(); d.K = t->getNumScenarios(); d.nodes = t->getNumNodes();
    d.n_nonleaf = t->getNumNonleafNodes(); d.n_children_tot = t->getNumChildrenTot();
    rn_tree tr{};
    tr.stages = t->getStageNodes(); tr.nodes_per_stage = t->getNodesPerStage();
    tr.nodes_per_stage_cumul = t->getNodesPerStageCumul(); tr.leaves = t->getLeaveArray(); tr.children = t->getChildArray();
    tr.ancestor = t->getAncestorArray(); tr.n_children = t->getNumChildren(); tr.n_children_cumul = t->getNumChildrenCumul();
    tr.prob = t->getProbArray(); tr.err_demand = t->getErrorDemandArray(); tr.err_price = t->getErrorPriceArray();
    rn_network nw{};
    nw.B = n->getMatB(); nw.Gd = n->getMatGd(); nw.E = n->getMatE(); nw.Ed = n->getMatEd();
    nw.xmin = n->getXmin(); nw.xmax = n->getXmax(); nw.xsafe = n->getXsafe(); nw.umin = n->getUmin(); nw.umax = n->getUmax();
    nw.alpha1 = n->getAlpha();
    rn_config c{};
    c.costW = smpcConfig->getCostW(); c.precond = smpcConfig->getMatPrcndDiag();
    c.penalty_x = smpcConfig->getPenaltyState(); c.penalty_xs = smpcConfig->getPenaltySafety();
    c.step_size = smpcConfig->getStepSize(); c.weight_economical = smpcConfig->getWeightEconomical();
    c.max_iterations = smpcConfig->getMaxIterations();
    int device = 0;
    if (const char *e = std::getenv("RAPIDNET_DEVICE")) device = std::atoi(e);
    rn_status rc = rn_create(&d, &tr, &nw, &c, device, &h);
    if (rc != RN_OK) {
        std::cerr << "rapidnet_b200: rn_create failed (" << rc << "): " << rn_last_error(nullptr) << std::endl;
        std::exit(EXIT_FAILURE);
    }
    // The reference has one formulation; the library's alternatives (include/rapidnet_b200.h: rn_sweep_mode,
    // rn_factor_mode -- same iterates up to fp32 rounding) are chosen from the environment so that the class API stays
    // the reference's: RAPIDNET_FACTORS = full | df | shared, RAPIDNET_SWEEP = persistent | chain | per_stage.
    rn_factor_mode fm = RN_FACTORS_FULL;
    rn_sweep_mode sm = RN_SWEEP_PERSISTENT;
    if (const char *e = std::getenv("RAPIDNET_FACTORS")) {
        const std::string v(e);
        if (v == "df") fm = RN_FACTORS_DF;
        else if (v == "shared") fm = RN_FACTORS_SHARED;
        else if (v != "full") { std::cerr << "rapidnet_b200: RAPIDNET_FACTORS=" << v << " is not full|df|shared" << std::endl; std::exit(EXIT_FAILURE); }
    }
    if (const char *e = std::getenv("RAPIDNET_SWEEP")) {
        const std::string v(e);
        if (v == "chain") sm = RN_SWEEP_CHAIN;
        else if (v == "per_stage") sm = RN_SWEEP_PER_STAGE;
        else if (v != "persistent") { std::cerr << "rapidnet_b200: RAPIDNET_SWEEP=" << v << " is not persistent|chain|per_stage" << std::endl; std::exit(EXIT_FAILURE); }
    }
    check(rn_set_modes(h, sm, fm), "rn_set_modes");
    // RAPIDNET_NULL_SPACE=config: use the null-space basis matL / matLhat that the controller configuration ships (the
    // reference parses them and then recomputes its own through cuSOLVER, Engine.cu:466-669; golden vectors produced with
    // another SVD -- the reference's engineTest.json comes from MATLAB -- only match in the basis of the file)
    if (const char *e = std::getenv("RAPIDNET_NULL_SPACE")) {
        if (std::string(e) == "config") check(rn_set_null_space(h, smpcConfig->getMatL(), smpcConfig->getMatLhat()), "rn_set_null_space");
        else if (std::string(e) != "svd") { std::cerr << "rapidnet_b200: RAPIDNET_NULL_SPACE=" << e << " is not config|svd" << std::endl; std::exit(EXIT_FAILURE); }
    }
    // RAPIDNET_GRID_LIMIT: share of the SMs for this Engine's persistent kernel when several controllers run side by side
    if (const char *e = std::getenv("RAPIDNET_GRID_LIMIT")) check(rn_set_grid_limit(h, std::atoi(e)), "rn_set_grid_limit");
}

// Engine.cuh:246.  The library does not use cuBLAS; a handle is created on first request for callers that expect the
// Engine to own one (the reference creates it in its constructor, Engine.cu:139).
cublasHandle_t Engine::getCublasHandle() {
    if (!cublasHandle && cublasCreate(&cublasHandle) != CUBLAS_STATUS_SUCCESS) {
        std::cerr << "rapidnet_b200: cublasCreate failed" << std::endl;
        std::exit(EXIT_FAILURE);
    }
    return cublasHandle;
}

// Ownership follows the reference: ~Engine releases the device memory and the cuBLAS handle but leaves the DwnNetwork and the
// ScenarioTree it created to the caller (Engine.cu:1431-1436 has those deletes commented out; the reference's tests delete the
// two objects themselves, Testing.cu:520-524, and would free them twice otherwise).
Engine::~Engine() {
    if (cublasHandle) cublasDestroy(cublasHandle);
    if (devMatGCopies) cudaFree(devMatGCopies);
    for (void *p : ownedDevice) cudaFree(p);
    if (h) rn_destroy(h);
}

void *Engine::toDevice(const void *host, size_t bytes) {
    void *d = nullptr;
    if (cudaMalloc(&d, bytes ? bytes : 1) != cudaSuccess || cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        std::cerr << "rapidnet_b200: device copy of " << bytes << " bytes failed: " << cudaGetErrorString(cudaGetLastError()) << std::endl;
        std::exit(EXIT_FAILURE);
    }
    ownedDevice.push_back(d);
    return d;
}

// Engine.cu:80-110, 191-230 (factor matrices), :336-360 (system matrices)
real_t **Engine::ptrTable(PtrTableId id) {
    if (devPtrTables[id]) return devPtrTables[id];
    ScenarioTree *t = ptrMyScenarioTree;
    const size_t nodes = t->getNumNodes(), N = t->getPredHorizon(), K = t->getNumScenarios();
    const size_t nx = ptrMySmpcConfig->getNX(), nu = ptrMySmpcConfig->getNU(), nv = ptrMySmpcConfig->getNV();
    const uint_t *nps = t->getNodesPerStage(), *cum = t->getNodesPerStageCumul();
    const size_t fb = t->getFinalBranchNode();
    std::vector<real_t *> tab(id == PT_G ? K : nodes, nullptr);
    auto per_node = [&](real_t *base, size_t stride) { for (size_t i = 0; i < tab.size(); i++) tab[i] = base + i * stride; };
    auto shared = [&](real_t *base) { for (size_t i = 0; i < tab.size(); i++) tab[i] = base; };
    auto aliased = [&](real_t *base, size_t stride) {   // Engine.cu:210-221
        for (size_t s = 0; s < N; s++)
            for (size_t j = 0; j < (size_t)nps[s]; j++) {
                const size_t c0 = (size_t)cum[s], cur = fb <= c0 ? fb - K + j : c0 + j;
                tab[c0 + j] = base + cur * stride;
            }
    };
    switch (id) {
        case PT_PHI: per_node(getMatPhi(), nv * 2 * nx); break;
        case PT_PSI: per_node(getMatPsi(), nv * nu); break;
        case PT_D: per_node(getMatD(), nv * 2 * nx); break;
        case PT_F: per_node(getMatF(), nv * nu); break;
        case PT_SIGMA: per_node(getMatSigma(), nv); break;
        case PT_OMEGA: aliased(getMatOmega(), nv * nv); break;
        case PT_THETA: aliased(getMatTheta(), nx * nv); break;
        case PT_G: per_node(getMatG(), nv * nx); break;
        case PT_SYS_B: shared(getSysMatB()); break;
        case PT_SYS_L: shared(getSysMatL()); break;
        case PT_SYS_LHAT: shared(getSysMatLhat()); break;
        case PT_SYS_F: per_node(getSysMatF(), 2 * nx * nx); break;
        case PT_SYS_G: per_node(getSysMatG(), nu * nu); break;
        default: break;
    }
    devPtrTables[id] = static_cast<real_t **>(toDevice(tab.data(), tab.size() * sizeof(real_t *)));
    return devPtrTables[id];
}

// Engine.cu:256-290: the host arrays of ScenarioTree, copied as they are
uint_t *Engine::treeU(int which) {
    if (devTreeU[which]) return devTreeU[which];
    ScenarioTree *t = ptrMyScenarioTree;
    const size_t nodes = t->getNumNodes(), N = t->getPredHorizon(), K = t->getNumScenarios(), nl = t->getNumNonleafNodes();
    const std::vector<uint_t> *src[7] = {&t->stageArray, &t->nodesPerStage, &t->nodesPerStageCumul, &t->leaveArray, &t->nChildArray,
                                         &t->ancestorArray, &t->nChildCumulArray};
    const size_t cnt[7] = {nodes, N + 1, N + 2, K, nl, nodes, nodes};   // the sizes the reference allocates (Engine.cu:256-262)
    std::vector<uint_t> host(cnt[which], 0);
    for (size_t i = 0; i < host.size() && i < src[which]->size(); i++) host[i] = (*src[which])[i];
    devTreeU[which] = static_cast<uint_t *>(toDevice(host.data(), host.size() * sizeof(uint_t)));
    return devTreeU[which];
}

real_t *Engine::treeF(int which) {
    if (devTreeF[which]) return devTreeF[which];
    ScenarioTree *t = ptrMyScenarioTree;
    const size_t nodes = t->getNumNodes(), nd = ptrMySmpcConfig->getND(), nu = ptrMySmpcConfig->getNU();
    const std::vector<real_t> *src[3] = {&t->probNodeArray, &t->errorDemandArray, &t->errorPriceArray};
    const size_t cnt[3] = {nodes, nodes * nd, nodes * nu};
    std::vector<real_t> host(cnt[which], 0.f);
    for (size_t i = 0; i < host.size() && i < src[which]->size(); i++) host[i] = (*src[which])[i];
    devTreeF[which] = static_cast<real_t *>(toDevice(host.data(), host.size() * sizeof(real_t)));
    return devTreeF[which];
}

real_t *Engine::buf(rn_buffer_id id) {
    void *p = nullptr;
    size_t bytes = 0;
    check(rn_buffer(h, id, &p, &bytes), "rn_buffer");
    return static_cast<real_t *>(p);
}

// The reference runs on the legacy default stream, so a cudaMemcpy issued by the caller right after any of these
// methods sees their results; the library's stream is non-blocking, hence the explicit rn_sync.
void Engine::factorStep() {
    check(rn_factor_step(h), "Engine::factorStep");
    check(rn_sync(h), "rn_sync");
    if (devMatGCopies) { cudaFree(devMatGCopies); devMatGCopies = nullptr; }   // G follows the null-space basis
    devPtrTables[PT_G] = nullptr;
}

// The library keeps ONE G = Bbar' (every scenario's copy is the same matrix, Engine.cu:736-747); callers of the reference index
// devMatG by scenario (its tests read copy `node - 1`, Testing.cu:437-440), so the getter hands out K copies, built on request.
real_t *Engine::getMatG() {
    if (devMatGCopies) return devMatGCopies;
    const size_t K = ptrMyScenarioTree->getNumScenarios(), sz = (size_t)ptrMySmpcConfig->getNV() * ptrMySmpcConfig->getNX();
    real_t *one = buf(RN_BUF_MAT_G);
    if (cudaMalloc(&devMatGCopies, K * sz * sizeof(real_t)) != cudaSuccess) { std::cerr << "rapidnet_b200: Engine::getMatG: out of device memory" << std::endl; std::exit(EXIT_FAILURE); }
    for (size_t k = 0; k < K; k++)
        if (cudaMemcpy(devMatGCopies + k * sz, one, sz * sizeof(real_t), cudaMemcpyDeviceToDevice) != cudaSuccess) { std::cerr << "rapidnet_b200: Engine::getMatG: copy failed" << std::endl; std::exit(EXIT_FAILURE); }
    return devMatGCopies;
}

void Engine::updateStateControl(real_t *currentX, real_t *prevU, real_t *prevDemand) {
    check(rn_update_state(h, currentX, prevU, prevDemand), "Engine::updateStateControl");
    check(rn_sync(h), "rn_sync");
}

void Engine::eliminateInputDistubanceCoupling(real_t *nominalDemand, real_t *nominalPrices) {
    check(rn_eliminate_coupling(h, nominalDemand, nominalPrices), "Engine::eliminateInputDistubanceCoupling");
    check(rn_sync(h), "rn_sync");
}

void Engine::setPriceUncertaintyFlag(bool inputFlag) {
    priceUncertaintyFlag = inputFlag;
    check(rn_set_uncertainty(h, demandUncertaintyFlag, priceUncertaintyFlag), "Engine::setPriceUncertaintyFlag");
}

void Engine::setDemandUncertaintyFlag(bool inputFlag) {
    demandUncertaintyFlag = inputFlag;
    check(rn_set_uncertainty(h, demandUncertaintyFlag, priceUncertaintyFlag), "Engine::setDemandUncertaintyFlag");
}

// ---- SmpcController ----------------------------------------------------------------------------------------------------
void SmpcController::check(rn_status rc, const char *what) {
    if (rc == RN_OK) return;
    std::cerr << "rapidnet_b200: " << what << " failed (" << rc << "): " << rn_last_error(ptrMyEngine->handle()) << std::endl;
    std::exit(EXIT_FAILURE);
}

void SmpcController::outOfScope(const char *what) {
    std::cerr << "rapidnet_b200: SmpcController::" << what << " belongs to the FBE / NAMA solvers of the reference, which this build "
                 "does not contain (controlAction always runs APG, SmpcController.cu:1617, 1646)" << std::endl;
    std::exit(EXIT_FAILURE);
}

void SmpcController::construct() {
    stepSize = ptrMySmpcConfig->getStepSize();
    factorStepFlag = false;
    simulatorFlag = true;
    vecPrimalInfs.assign((size_t)ptrMySmpcConfig->getMaxIterations() + 1, 0.f);
    economicKpi = smoothKpi = safeKpi = networkKpi = 0;
    if (!ptrMyEngine->getApgFlag())
        std::cerr << "rapidnet_b200: algorithmName selects FBE/NAMA buffers in the reference; controlAction always runs APG "
                     "(SmpcController.cu:1617, 1646) and so does this build" << std::endl;
    refreshDevicePointers();
}

SmpcController::SmpcController(Forecaster *myForecaster, Engine *myEngine, SmpcConfiguration *mySmpcConfig) {
    ptrMyForecaster = myForecaster;
    ptrMyEngine = myEngine;
    ptrMySmpcConfig = mySmpcConfig;
    construct();
}

SmpcController::SmpcController(std::string pathToConfigFile) {
    ptrMySmpcConfig = new SmpcConfiguration(pathToConfigFile);
    ptrMyForecaster = new Forecaster(ptrMySmpcConfig->getPathToForecaster());
    ptrMyEngine = new Engine(ptrMySmpcConfig);
    ownsObjects = true;   // released by releaseObjects(), or by the caller as in the reference's tests
    construct();
}

// Like the reference (SmpcController.cu:2091-2102 frees its own buffers only), the controller does not delete the configuration,
// the forecaster and the engine -- not even the ones SmpcController(string) created (:78-82): its callers do
// (Testing.cu:520-526 deletes every one of them BEFORE the controller; nothing here may touch them).
// releaseObjects() is the one-call version of those deletes for callers written against this library.
SmpcController::~SmpcController() {}

void SmpcController::releaseObjects() {
    if (!ownsObjects) return;
    DwnNetwork *n = ptrMyEngine ? ptrMyEngine->getDwnNetwork() : nullptr;
    ScenarioTree *t = ptrMyEngine ? ptrMyEngine->getScenarioTree() : nullptr;
    delete ptrMyEngine; delete n; delete t;
    delete ptrMyForecaster;
    delete ptrMySmpcConfig;
    ptrMyEngine = nullptr; ptrMyForecaster = nullptr; ptrMySmpcConfig = nullptr;
    ownsObjects = false;
}

void SmpcController::refreshDevicePointers() {
    rn_handle *h = ptrMyEngine->handle();
    auto get = [&](rn_buffer_id id) {
        void *p = nullptr;
        size_t bytes = 0;
        check(rn_buffer(h, id, &p, &bytes), "rn_buffer");
        return static_cast<real_t *>(p);
    };
    devVecX = get(RN_BUF_VEC_X); devVecU = get(RN_BUF_VEC_U); devVecV = get(RN_BUF_VEC_V);
    devVecXi = get(RN_BUF_VEC_XI); devVecPsi = get(RN_BUF_VEC_PSI);
    devVecAcceleratedXi = get(RN_BUF_VEC_ACCEL_XI); devVecAcceleratedPsi = get(RN_BUF_VEC_ACCEL_PSI);
    devVecPrimalXi = get(RN_BUF_VEC_PRIMAL_XI); devVecPrimalPsi = get(RN_BUF_VEC_PRIMAL_PSI);
    devVecDualXi = get(RN_BUF_VEC_DUAL_XI); devVecDualPsi = get(RN_BUF_VEC_DUAL_PSI);
    devVecUpdateXi = get(RN_BUF_VEC_UPDATE_XI); devVecUpdatePsi = get(RN_BUF_VEC_UPDATE_PSI);
    devVecFixedPointResidualXi = get(RN_BUF_VEC_RESIDUAL_XI); devVecFixedPointResidualPsi = get(RN_BUF_VEC_RESIDUAL_PSI);
    devControlAction = get(RN_BUF_CONTROL_ACTION); devStateUpdate = get(RN_BUF_STATE_UPDATE);
    proximalSlots[0] = devVecAcceleratedXi; proximalSlots[1] = devVecAcceleratedPsi;   // SmpcController.cu:510-511
    ptrProximalXi = &proximalSlots[0]; ptrProximalPsi = &proximalSlots[1];
}

void SmpcController::initialiseSmpcController() {
    factorStepFlag = true;
    ptrMyEngine->factorStep();
    ptrMyEngine->updateStateControl(ptrMySmpcConfig->getCurrentX(), ptrMySmpcConfig->getPrevU(), ptrMySmpcConfig->getPrevDemand());
    ptrMyEngine->eliminateInputDistubanceCoupling(ptrMyForecaster->getNominalDemand(), ptrMyForecaster->getNominalPrices());
    refreshDevicePointers();
}

void SmpcController::initialiseAlgorithm() {
    check(rn_apg_init(ptrMyEngine->handle()), "SmpcController::initialiseAlgorithm");
    check(rn_sync(ptrMyEngine->handle()), "rn_sync");
    refreshDevicePointers();
}

// one protected step, synchronous like the reference's (see Engine::factorStep)
void SmpcController::step(rn_step_kind kind, real_t lambda, const char *what) {
    check(rn_step(ptrMyEngine->handle(), kind, lambda), what);
    check(rn_sync(ptrMyEngine->handle()), "rn_sync");
}
void SmpcController::dualExtrapolationStep(real_t lambda) { step(RN_STEP_EXTRAPOLATE, lambda, "SmpcController::dualExtrapolationStep"); }
// The reference factors lazily: solveStep() runs initialiseSmpcController() when no factor step has happened yet
// (SmpcController.cu:579-582), so a fresh controller may go straight to controlAction / algorithmApg.  Every entry that
// reaches solveStep in the reference does the same here, before the C ABI (which insists on the order) is called.
void SmpcController::ensureFactored() {
    if (!factorStepFlag) initialiseSmpcController();
}
void SmpcController::solveStep() { ensureFactored(); step(RN_STEP_SOLVE, 0.f, "SmpcController::solveStep"); }
void SmpcController::proximalFunG() { step(RN_STEP_PROX, 0.f, "SmpcController::proximalFunG"); }
void SmpcController::computeFixedPointResidual() { step(RN_STEP_RESIDUAL, 0.f, "SmpcController::computeFixedPointResidual"); }
void SmpcController::dualUpdate() { step(RN_STEP_DUAL_UPDATE, 0.f, "SmpcController::dualUpdate"); }

uint_t SmpcController::algorithmApg() {
    ensureFactored();
    const uint_t iters = ptrMySmpcConfig->getMaxIterations();
    check(rn_apg_solve(ptrMyEngine->handle(), iters, nullptr, vecPrimalInfs.data()), "SmpcController::algorithmApg");
    refreshDevicePointers();
    return 1;
}

void SmpcController::controllerSmpc() {
    ensureFactored();
    ptrMyEngine->updateStateControl(ptrMySmpcConfig->getCurrentX(), ptrMySmpcConfig->getPrevU(), ptrMySmpcConfig->getPrevDemand());
    ptrMyEngine->eliminateInputDistubanceCoupling(ptrMyForecaster->getNominalDemand(), ptrMyForecaster->getNominalPrices());
    algorithmApg();
}

uint_t SmpcController::controlAction(real_t *u) {
    ensureFactored();
    check(rn_control_action(ptrMyEngine->handle(), ptrMySmpcConfig->getCurrentX(), ptrMySmpcConfig->getPrevU(),
                            ptrMySmpcConfig->getPrevDemand(), ptrMyForecaster->getNominalDemand(),
                            ptrMyForecaster->getNominalPrices(), ptrMySmpcConfig->getMaxIterations(), /*clamp=*/0, u),
          "SmpcController::controlAction");
    refreshDevicePointers();
    return 1;   // the reference returns 0 only when cudaMemGetInfo sees a leak (:1619-1624); nothing is allocated here
}

uint_t SmpcController::controlAction(std::fstream &controlOutputJson) {
    if (!controlOutputJson.is_open()) return 0;
    ensureFactored();
    const uint_t nu = ptrMySmpcConfig->getNU();
    std::vector<real_t> currentControl(nu);
    check(rn_control_action(ptrMyEngine->handle(), ptrMySmpcConfig->getCurrentX(), ptrMySmpcConfig->getPrevU(),
                            ptrMySmpcConfig->getPrevDemand(), ptrMyForecaster->getNominalDemand(),
                            ptrMyForecaster->getNominalPrices(), ptrMySmpcConfig->getMaxIterations(), /*clamp=*/1,
                            currentControl.data()),
          "SmpcController::controlAction");
    refreshDevicePointers();
    controlOutputJson << "\"control\" : [";   // same pseudo-JSON fragment as the reference (:1651-1658)
    for (uint_t i = 0; i < nu; i++) controlOutputJson << currentControl[i] << ", ";
    controlOutputJson << "]" << std::endl;
    return 1;
}

void SmpcController::moveForewardInTime() {
    if (simulatorFlag) {
        const uint_t nx = ptrMySmpcConfig->getNX(), nu = ptrMySmpcConfig->getNU();
        std::vector<real_t> stateUpdate(nx), previousControl(nu);
        check(rn_move_forward(ptrMyEngine->handle(), stateUpdate.data(), previousControl.data()), "SmpcController::moveForewardInTime");
        updateKpi(stateUpdate.data(), previousControl.data());
        ptrMySmpcConfig->setCurrentState(stateUpdate.data());
        ptrMySmpcConfig->setPreviousControl(previousControl.data());
        ptrMySmpcConfig->setpreviousdemand(ptrMyForecaster->getNominalDemand());
    } else {
        ptrMySmpcConfig->setCurrentState();
        ptrMySmpcConfig->setPreviousControl();
        ptrMySmpcConfig->setPreviousDemand();
    }
}

// Stand-alone form of the infeasibility measure (SmpcController.cu:1480-1496): signed residual at the arg-max-abs of each
// block, the larger of the two.  Inside algorithmApg the persistent kernel logs the same quantity per iteration
// (vecPrimalInfs); this one serves callers that drive the steps by hand, with the same two cuBLAS calls as the reference.
real_t SmpcController::updatePrimalInfeasibity() {
    refreshDevicePointers();
    check(rn_sync(ptrMyEngine->handle()), "rn_sync");
    const int nx = (int)ptrMyEngine->getDwnNetwork()->getNumTanks(), nu = (int)ptrMyEngine->getDwnNetwork()->getNumControls();
    const int nodes = (int)ptrMyEngine->getScenarioTree()->getNumNodes();
    cublasHandle_t cb = ptrMyEngine->getCublasHandle();
    int ixi = 0, ipsi = 0;
    real_t vxi = 0, vpsi = 0;
    if (cublasIsamax(cb, 2 * nx * nodes, devVecFixedPointResidualXi, 1, &ixi) != CUBLAS_STATUS_SUCCESS ||
        cublasIsamax(cb, nu * nodes, devVecFixedPointResidualPsi, 1, &ipsi) != CUBLAS_STATUS_SUCCESS ||
        cudaMemcpy(&vxi, devVecFixedPointResidualXi + (ixi - 1), sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(&vpsi, devVecFixedPointResidualPsi + (ipsi - 1), sizeof(real_t), cudaMemcpyDeviceToHost) != cudaSuccess) {
        std::cerr << "rapidnet_b200: SmpcController::updatePrimalInfeasibity failed" << std::endl;
        std::exit(EXIT_FAILURE);
    }
    return std::max(vxi, vpsi);
}

void SmpcController::updateKpi(real_t *state, real_t *control) {
    const uint_t nx = ptrMySmpcConfig->getNX(), nu = ptrMySmpcConfig->getNU();
    real_t *safeX = ptrMyEngine->getDwnNetwork()->getXsafe();
    real_t *constantPrice = ptrMyEngine->getDwnNetwork()->getAlpha();
    real_t *variablePrice = ptrMyForecaster->getNominalPrices();
    real_t *previousControl = ptrMySmpcConfig->getPrevU();
    const real_t weightEconomic = ptrMySmpcConfig->getWeightEconomical();
    real_t ecoKpi = 0, smKpi = 0, saKpi = 0, netKpi = 0;
    for (uint_t i = 0; i < nu; i++) {
        ecoKpi = ecoKpi + weightEconomic * (constantPrice[i] + variablePrice[i]) * std::fabs(control[i]);
        const real_t deltaU = previousControl[i] - control[i];
        smKpi = smKpi + deltaU * deltaU;
    }
    for (uint_t i = 0; i < nx; i++) {
        real_t waterLevel = state[i] - safeX[i];
        if (waterLevel > 0) waterLevel = 0;
        saKpi = saKpi + std::fabs(waterLevel);
        netKpi = netKpi + std::fabs(state[i]);
    }
    economicKpi += ecoKpi; smoothKpi += smKpi; safeKpi += saKpi; networkKpi += netKpi;
}

real_t SmpcController::getEconomicKpi(uint_t simulationTime) { return economicKpi / 3600 / simulationTime; }
real_t SmpcController::getSmoothKpi(uint_t simulationTime) { return smoothKpi / 3600 / simulationTime; }
real_t SmpcController::getNetworkKpi(uint_t simulationTime) {
    real_t safeLevelNorm = 0;
    const uint_t nx = ptrMySmpcConfig->getNX();
    for (uint_t i = 0; i < nx; i++) safeLevelNorm += getDwnNetwork()->getXsafe()[i];
    return 100 * simulationTime * safeLevelNorm / networkKpi;
}
real_t SmpcController::getSafetyKpi(uint_t) { return safeKpi; }

}  // namespace rapidnet
