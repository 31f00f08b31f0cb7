// rapidnet_host.hpp -- C++ host side of rapidnet-b200: the reference's class surface over the C ABI.
//
// GPUEngineering/RapidNet has no plugin interface: callers (its main() and its tests) use six C++ classes.  These are
// the same classes -- same names, same getters, same argument meaning, same error behaviour (print + exit; ScenarioTree
// throws) -- implemented as thin host code over include/rapidnet_b200.h:
//
//   DwnNetwork          /root/reference/src/DwnNetwork.cuh:67-147          JSON -> host arrays
//   ScenarioTree        /root/reference/src/ScenarioTree.cuh:64-158        JSON -> host arrays
//   Forecaster          /root/reference/src/Forecaster.cuh:57-95           JSON -> nominal demand / prices per time slot
//   SmpcConfiguration   /root/reference/src/SmpcConfiguration.cuh:60-185   JSON -> controller configuration
//   Engine              /root/reference/src/Engine.cuh:58-367              device memory owner, factor step, affine terms
//   SmpcController      /root/reference/src/SmpcController.cuh:50-462      APG loop, controlAction, closed-loop bookkeeping
//
// Engine owns the rn_handle; every device-pointer getter hands out a borrowed pointer into the library's buffers in
// the reference's packed layouts (valid until ~Engine).  The protected step methods and device-buffer members of
// SmpcController exist because the reference's tests subclass the controller and poke them
// (/root/reference/src/test/TestSmpcController.cu:134-398).
#pragma once

#include <fstream>
#include <string>
#include <vector>

#include "../../include/rapidnet_b200.h"

// the declaration of cublas_api.h, repeated so that this header does not pull in cuBLAS (Engine::getCublasHandle)
struct cublasContext;
typedef struct cublasContext *cublasHandle_t;

namespace rapidnet {

typedef float real_t;   // /root/reference/src/Configuration.h:30-31
typedef int uint_t;

class DwnNetwork {
public:
    explicit DwnNetwork(std::string pathToFile);
    uint_t getNumTanks() { return nTanks; }
    uint_t getNumControls() { return nControl; }
    uint_t getNumDemands() { return nDemand; }
    uint_t getNumMixNodes() { return nMixNodes; }
    real_t *getMatA() { return matA.data(); }
    real_t *getMatB() { return matB.data(); }
    real_t *getMatGd() { return matGd.data(); }
    real_t *getMatE() { return matE.data(); }
    real_t *getMatEd() { return matEd.data(); }
    real_t *getXmin() { return vecXmin.data(); }
    real_t *getXmax() { return vecXmax.data(); }
    real_t *getXsafe() { return vecXsafe.data(); }
    real_t *getUmin() { return vecUmin.data(); }
    real_t *getUmax() { return vecUmax.data(); }
    real_t *getAlpha() { return vecCostAlpha1.data(); }

private:
    uint_t nTanks = 0, nControl = 0, nDemand = 0, nMixNodes = 0;
    std::vector<real_t> matA, matB, matGd, matE, matEd, vecXmin, vecXmax, vecXsafe, vecUmin, vecUmax, vecCostAlpha1;
};

class ScenarioTree {
public:
    explicit ScenarioTree(std::string pathToFile);   // throws std::logic_error when the file is missing (ScenarioTree.cu:40)
    uint_t getPredHorizon() { return nPredHorizon; }
    uint_t getNumScenarios() { return nScenario; }
    uint_t getNumNodes() { return nNodes; }
    uint_t getNumChildrenTot() { return nChildrenTot; }
    uint_t getNumNonleafNodes() { return nNonLeafNodes; }
    uint_t getFinalBranchNode();    // ScenarioTree.cu:149-158
    uint_t getFinalBranchStage();   // ScenarioTree.cu:160-169
    uint_t *getStageNodes() { return stageArray.data(); }
    uint_t *getNodesPerStage() { return nodesPerStage.data(); }
    uint_t *getNodesPerStageCumul() { return nodesPerStageCumul.data(); }
    uint_t *getLeaveArray() { return leaveArray.data(); }
    uint_t *getChildArray() { return childArray.data(); }
    uint_t *getAncestorArray() { return ancestorArray.data(); }
    uint_t *getNumChildren() { return nChildArray.data(); }
    uint_t *getNumChildrenCumul() { return nChildCumulArray.data(); }
    real_t *getProbArray() { return probNodeArray.data(); }
    real_t *getErrorDemandArray() { return errorDemandArray.data(); }
    real_t *getErrorPriceArray() { return errorPriceArray.data(); }

private:
    friend class Engine;   // device copies of the arrays (Engine::getTree*)
    uint_t nPredHorizon = 0, nScenario = 0, nNodes = 0, nChildrenTot = 0, nNonLeafNodes = 0;
    std::vector<uint_t> stageArray, nodesPerStage, nodesPerStageCumul, leaveArray, childArray, ancestorArray, nChildArray,
        nChildCumulArray;
    std::vector<real_t> probNodeArray, errorDemandArray, errorPriceArray;
};

class Forecaster {
public:
    explicit Forecaster(std::string pathToFile);
    virtual ~Forecaster() {}
    uint_t getPredHorizon() { return nPredHorizon; }
    uint_t getSimHorizon() { return simHorizon; }
    uint_t getDimDemand() { return dimDemand; }
    uint_t getDimPrice() { return dimPrices; }
    real_t *getNominalDemand() { return nominalDemand.data(); }
    real_t *getNominalPrices() { return nominalPrice.data(); }
    // member 4 + 2 t / 5 + 2 t of the JSON object, by POSITION (Forecaster.cu:93-119); 1 = found, 0 = past the end
    virtual uint_t predictDemand(uint_t simTime);
    virtual uint_t predictPrices(uint_t simTime);

private:
    uint_t nPredHorizon = 0, simHorizon = 0, dimDemand = 0, dimPrices = 0;
    std::vector<real_t> nominalDemand, nominalPrice;
    std::vector<std::vector<real_t>> slots;   // positional members 4, 5, 6, ... of the file
};

class SmpcConfiguration {
public:
    explicit SmpcConfiguration(std::string pathToFile);
    uint_t getNX() { return NX; }
    uint_t getNU() { return NU; }
    uint_t getND() { return ND; }
    uint_t getNV() { return NV; }
    uint_t getLbfgsBufferSize() { return lbfgsBufferSize; }
    real_t *getMatL() { return matL.data(); }
    real_t *getMatLhat() { return matLhat.data(); }
    real_t *getMatPrcndDiag() { return matDiagPrecnd.data(); }
    real_t *getCostW() { return matCostW.data(); }
    real_t *getCurrentX() { return currentX.data(); }
    real_t *getPrevU() { return prevU.data(); }
    real_t *getPrevDemand() { return prevDemand.data(); }
    real_t getPenaltyState() { return penaltyStateX; }
    real_t getPenaltySafety() { return penaltySafetyX; }
    uint_t getMaxIterations() { return maxIteration; }
    real_t getStepSize() { return stepSize; }
    std::string getPathToControllerConfig() { return pathToConfiguration; }
    std::string getPathToNetwork() { return pathToNetwork; }
    std::string getPathToScenarioTree() { return pathToScenarioTree; }
    std::string getPathToForecaster() { return pathToForecaster; }
    real_t getWeightEconomical() { return weightPrice; }
    std::string getOptimisationAlgorithm() { return algorithmName; }
    void setCurrentState();      // re-read currentX from the configuration file (SmpcConfiguration.cu:241-256)
    void setPreviousControl();   // re-read prevU
    void setPreviousDemand();    // re-read prevDemand -- into prevU, as the reference does (:290)
    void setCurrentState(real_t *state);
    void setPreviousControl(real_t *control);
    void setpreviousdemand(real_t *demand);

private:
    uint_t NX = 0, NU = 0, ND = 0, NV = 0, N = 0, maxIteration = 0, lbfgsBufferSize = 0;
    std::vector<real_t> matL, matLhat, matCostW, matDiagPrecnd, currentX, prevU, prevDemand;
    real_t penaltyStateX = 0, penaltySafetyX = 0, stepSize = 0;
    real_t weightPrice = 1, weightSmooth = 1, weightSafety = 1;   // hard-coded (SmpcConfiguration.cu:38-40)
    std::string pathToConfiguration, pathToNetwork, pathToScenarioTree, pathToForecaster, algorithmName;
};

class Engine {
public:
    explicit Engine(SmpcConfiguration *smpcConfig);   // loads network + tree from the configuration's paths (Engine.cu:126-163)
    ~Engine();
    void factorStep();                                                            // Engine.cu:671-774
    void updateStateControl(real_t *currentX, real_t *prevU, real_t *prevDemand); // Engine.cu:1300-1316
    void eliminateInputDistubanceCoupling(real_t *nominalDemand, real_t *nominalPrices);   // Engine.cu:1147-1298
    ScenarioTree *getScenarioTree() { return ptrMyScenarioTree; }
    DwnNetwork *getDwnNetwork() { return ptrMyNetwork; }
    // device pointers, reference layouts (Engine.cuh:104-318)
    real_t *getSysMatB() { return buf(RN_BUF_SYS_MAT_B); }
    real_t *getSysMatF() { return buf(RN_BUF_SYS_MAT_F); }
    real_t *getSysMatG() { return buf(RN_BUF_SYS_MAT_G); }
    real_t *getSysMatL() { return buf(RN_BUF_SYS_MAT_L); }
    real_t *getSysMatLhat() { return buf(RN_BUF_SYS_MAT_LHAT); }
    real_t *getVecPreviousControl() { return buf(RN_BUF_VEC_PREV_CONTROL); }
    real_t *getVecCurrentState() { return buf(RN_BUF_VEC_CURRENT_STATE); }
    real_t *getVecPreviousUhat() { return buf(RN_BUF_VEC_PREV_UHAT); }
    real_t *getVecDemand() { return buf(RN_BUF_VEC_PREV_DEMAND); }
    real_t *getMatPhi() { return buf(RN_BUF_MAT_PHI); }
    real_t *getMatPsi() { return buf(RN_BUF_MAT_PSI); }
    real_t *getMatTheta() { return buf(RN_BUF_MAT_THETA); }
    real_t *getMatOmega() { return buf(RN_BUF_MAT_OMEGA); }
    real_t *getMatSigma() { return buf(RN_BUF_MAT_SIGMA); }
    real_t *getMatD() { return buf(RN_BUF_MAT_D); }
    real_t *getMatF() { return buf(RN_BUF_MAT_F); }
    real_t *getMatG();   // K identical copies of G = Bbar', as the reference lays it out (Engine.cu:173, 204-207: one per scenario)
    real_t *getVecUhat() { return buf(RN_BUF_VEC_UHAT); }
    real_t *getVecBeta() { return buf(RN_BUF_VEC_BETA); }
    real_t *getVecE() { return buf(RN_BUF_VEC_E); }
    real_t *getSysXmin() { return buf(RN_BUF_SYS_XMIN); }
    real_t *getSysXmax() { return buf(RN_BUF_SYS_XMAX); }
    real_t *getSysXs() { return buf(RN_BUF_SYS_XS); }
    real_t *getSysXsUpper() { return buf(RN_BUF_SYS_XS_UPPER); }
    real_t *getSysUmin() { return buf(RN_BUF_SYS_UMIN); }
    real_t *getSysUmax() { return buf(RN_BUF_SYS_UMAX); }
    real_t *getPriceAlpha() { return buf(RN_BUF_VEC_ALPHA); }
    bool getPriceUncertainty() { return priceUncertaintyFlag; }
    bool getDemandUncertantiy() { return demandUncertaintyFlag; }
    bool getApgFlag() { return apgFlag; }
    bool getGlobalFbeFlag() { return globalFbeFlag; }
    bool getNamaFlag() { return namaFlag; }
    void setPriceUncertaintyFlag(bool inputFlag);
    void setDemandUncertaintyFlag(bool inputFlag);
    // Per-node pointer tables on the device (Engine.cuh:128-230; built by Engine.cu:80-110, 191-230, 336-360): entry i points at
    // node i's matrix inside the packed arrays above.  The library itself does not use them (its kernels index the packed
    // arrays); they are built on first use for callers that feed cuBLAS batched routines the way the reference does.
    // Omega / Theta alias from the final branching node on (Engine.cu:210-221); B, L, Lhat and G exist once here (the
    // reference keeps one copy per scenario), so every entry of those tables points at the same matrix.
    real_t **getPtrSysMatB() { return ptrTable(PT_SYS_B); }
    real_t **getPtrSysMatF() { return ptrTable(PT_SYS_F); }
    real_t **getPtrSysMatG() { return ptrTable(PT_SYS_G); }
    real_t **getPtrSysMatL() { return ptrTable(PT_SYS_L); }
    real_t **getPtrSysMatLhat() { return ptrTable(PT_SYS_LHAT); }
    real_t **getPtrMatPhi() { return ptrTable(PT_PHI); }
    real_t **getPtrMatPsi() { return ptrTable(PT_PSI); }
    real_t **getPtrMatTheta() { return ptrTable(PT_THETA); }
    real_t **getPtrMatOmega() { return ptrTable(PT_OMEGA); }
    real_t **getPtrMatSigma() { return ptrTable(PT_SIGMA); }
    real_t **getPtrMatD() { return ptrTable(PT_D); }
    real_t **getPtrMatF() { return ptrTable(PT_F); }
    real_t **getPtrMatG() { return ptrTable(PT_G); }
    // The scenario tree on the device, the host arrays of ScenarioTree verbatim (Engine.cu:256-290; 1-based node ids as in the
    // JSON).  Also built on first use: the library keeps its own 0-based copy.
    uint_t *getTreeStages() { return treeU(0); }
    uint_t *getTreeNodesPerStage() { return treeU(1); }
    uint_t *getTreeNodesPerStageCumul() { return treeU(2); }
    uint_t *getTreeLeaves() { return treeU(3); }
    uint_t *getTreeNumChildren() { return treeU(4); }
    uint_t *getTreeAncestor() { return treeU(5); }
    uint_t *getTreeNumChildrenCumul() { return treeU(6); }
    real_t *getTreeProb() { return treeF(0); }
    real_t *getTreeErrorDemand() { return treeF(1); }
    real_t *getTreeErrorPrices() { return treeF(2); }
    cublasHandle_t getCublasHandle();
    // the C-ABI handle underneath (what a binding of the reference would hold)
    rn_handle *handle() { return h; }

private:
    enum PtrTableId { PT_SYS_B, PT_SYS_F, PT_SYS_G, PT_SYS_L, PT_SYS_LHAT, PT_PHI, PT_PSI, PT_THETA, PT_OMEGA, PT_SIGMA, PT_D, PT_F,
                      PT_G, PT_COUNT_ };
    real_t **ptrTable(PtrTableId id);
    uint_t *treeU(int which);
    real_t *treeF(int which);
    void *toDevice(const void *host, size_t bytes);
    real_t *devMatGCopies = nullptr;   // getMatG(): rebuilt after every factor step
    real_t **devPtrTables[PT_COUNT_] = {};
    uint_t *devTreeU[7] = {};
    real_t *devTreeF[3] = {};
    std::vector<void *> ownedDevice;
    cublasHandle_t cublasHandle = nullptr;
    real_t *buf(rn_buffer_id id);
    void check(rn_status rc, const char *what);
    DwnNetwork *ptrMyNetwork = nullptr;
    ScenarioTree *ptrMyScenarioTree = nullptr;
    SmpcConfiguration *ptrMySmpcConfig = nullptr;
    rn_handle *h = nullptr;
    bool priceUncertaintyFlag = true, demandUncertaintyFlag = true;
    bool apgFlag = true, globalFbeFlag = false, namaFlag = false;
};

class SmpcController {
public:
    SmpcController(Forecaster *myForecaster, Engine *myEngine, SmpcConfiguration *mySmpcConfig);
    explicit SmpcController(std::string pathToConfigFile);
    virtual ~SmpcController();
    void releaseObjects();   // not in the reference: deletes what SmpcController(string) created (the reference leaves that to its callers)
    void initialiseSmpcController();                       // SmpcController.cu:476-487
    void controllerSmpc();                                 // :1593-1599
    uint_t controlAction(real_t *u);                       // :1607-1625   1 = success
    uint_t controlAction(std::fstream &controlOutputJson); // :1633-1667
    DwnNetwork *getDwnNetwork() { return ptrMyEngine->getDwnNetwork(); }
    ScenarioTree *getScenarioTree() { return ptrMyEngine->getScenarioTree(); }
    SmpcConfiguration *getSmpcConfiguration() { return ptrMySmpcConfig; }
    Forecaster *getForecaster() { return ptrMyForecaster; }
    Engine *getEngine() { return ptrMyEngine; }
    void moveForewardInTime();                             // :1679-1717
    real_t getEconomicKpi(uint_t simulationTime);          // :1824-1859
    real_t getSmoothKpi(uint_t simulationTime);
    real_t getNetworkKpi(uint_t simulationTime);
    real_t getSafetyKpi(uint_t simulationTime);
    void updateKpi(real_t *state, real_t *control);        // :1778-1818
    real_t *getPrimalInfeasibility() { return vecPrimalInfs.data(); }

protected:
    void dualExtrapolationStep(real_t lambda);             // :535-557
    void solveStep();                                      // :563-755
    void proximalFunG();                                   // :759-835
    void computeFixedPointResidual();                      // :839-850
    void dualUpdate();                                     // :854-864
    uint_t algorithmApg();                                 // :1500-1525
    real_t updatePrimalInfeasibity();                      // :1480-1496 (stand-alone; inside the APG loop the kernel logs it)
    void initialiseAlgorithm();                            // :420-450
    void ensureFactored();                                 // the lazy factor step of solveStep (:579-582)
    // the library's device buffers under the reference's member names; xi / psi / update roles swap physical
    // buffers with the iteration parity, so the pointers are refreshed after every call that runs iterations
    void refreshDevicePointers();
    real_t *devVecX = nullptr, *devVecU = nullptr, *devVecV = nullptr;
    real_t *devVecXi = nullptr, *devVecPsi = nullptr, *devVecAcceleratedXi = nullptr, *devVecAcceleratedPsi = nullptr;
    real_t *devVecPrimalXi = nullptr, *devVecPrimalPsi = nullptr, *devVecDualXi = nullptr, *devVecDualPsi = nullptr;
    real_t *devVecUpdateXi = nullptr, *devVecUpdatePsi = nullptr;
    real_t *devVecFixedPointResidualXi = nullptr, *devVecFixedPointResidualPsi = nullptr;
    real_t *devControlAction = nullptr, *devStateUpdate = nullptr;
    // where the prox reads its argument (SmpcController.cuh:429-433): entry 0 is the accelerated dual w in APG mode (:510-511)
    real_t **ptrProximalXi = nullptr, **ptrProximalPsi = nullptr;
    // ---- members of the solvers that are OUT OF SCOPE here (global FBE, NAMA, L-BFGS: SmpcController.cuh:203-300, 417-627;
    // never reached from controlAction).  They are declared so that code written against the
    // reference's class -- its own TestSmpcController among it -- compiles; the buffers stay null and the methods stop the
    // program with a message, like every error of the reference does.
    void computeHessianOracalGlobalFbe() { outOfScope("computeHessianOracalGlobalFbe"); }
    void computeGradientFbe() { outOfScope("computeGradientFbe"); }
    void computeLbfgsDirection() { outOfScope("computeLbfgsDirection"); }
    real_t computeLineSearchLbfgsUpdate(real_t) { outOfScope("computeLineSearchLbfgsUpdate"); return 0; }
    real_t computeLineSearchAmeLbfgsUpdate(real_t) { outOfScope("computeLineSearchAmeLbfgsUpdate"); return 0; }
    real_t computeValueFbe() { outOfScope("computeValueFbe"); return 0; }
    void updateFixedPointResidualNamaAlgorithm() { outOfScope("updateFixedPointResidualNamaAlgorithm"); }
    real_t *devVecResidual = nullptr, *devVecGradientFbeXi = nullptr, *devVecGradientFbePsi = nullptr;
    real_t **devPtrVecHessianOracleXi = nullptr, **devPtrVecHessianOraclePsi = nullptr;
    real_t *devVecXdir = nullptr, *devVecUdir = nullptr, *devVecPrevXi = nullptr, *devVecPrevPsi = nullptr;
    real_t *devVecLbfgsDirXi = nullptr, *devVecLbfgsDirPsi = nullptr;
    real_t **ptrLbfgsCurrentYvecXi = nullptr, **ptrLbfgsCurrentYvecPsi = nullptr, **ptrLbfgsPreviousYvecXi = nullptr,
           **ptrLbfgsPreviousYvecPsi = nullptr;
    real_t *devLbfgsBufferMatS = nullptr, *devLbfgsBufferMatY = nullptr, *lbfgsBufferRho = nullptr;
    uint_t lbfgsBufferCol = 0, lbfgsBufferMemory = 0;
    real_t lbfgsBufferHessian = 0;
    Forecaster *ptrMyForecaster = nullptr;
    Engine *ptrMyEngine = nullptr;
    SmpcConfiguration *ptrMySmpcConfig = nullptr;
    std::vector<real_t> vecPrimalInfs;
    real_t stepSize = 0;
    bool factorStepFlag = false, simulatorFlag = true;
    real_t economicKpi = 0, smoothKpi = 0, safeKpi = 0, networkKpi = 0;

private:
    void construct();
    void check(rn_status rc, const char *what);
    void step(rn_step_kind kind, real_t lambda, const char *what);
    [[noreturn]] void outOfScope(const char *what);
    real_t *proximalSlots[2] = {nullptr, nullptr};
    bool ownsObjects = false;
};

}  // namespace rapidnet
