// host_tests.cpp -- the reference's own test suite replayed against the C++ facade (rapidnet_host.hpp).
//
// Mirrors /root/reference/src/test/Testing.cu (loader tests :78-335, testEngineTesting :340-477, testSmpcController
// :482-531) and /root/reference/src/test/TestSmpcController.cu:114-398 (a subclass of SmpcController that pokes golden
// inputs into the protected device buffers and calls the protected step methods), with the reference's tolerances
// (Testing.cu:33-76: abs 1e-2; TestSmpcController.cu:28-47: abs 0.1, or 0.1 % relative when |value| > 100).
//
//   host_tests loaders    <controllerConfig.json>                         (no GPU needed)
//   host_tests engine     <controllerConfig.json> <engineTest.json>
//   host_tests smpc       <controllerConfig.json> <engineTest.json> <smpcTest.json>
//   host_tests closedloop <controllerConfig.json> <steps> <out.json>      (main.cu:27-69 as a function)
// Exit code 0 = all assertions hold; the first failing one prints file:line and exits 1 (like _ASSERT).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <chrono>
#include <thread>
#include <vector>

#include "rapidjson/document.h"
#include "rapidjson/filereadstream.h"

#include "rapidnet_host.hpp"

using namespace rapidnet;

#define T_ASSERT(c)                                                                           \
    do {                                                                                      \
        if (!(c)) { std::cerr << "ASSERTION FAILED " << __FILE__ << ":" << __LINE__ << "  " #c << std::endl; std::exit(1); } \
    } while (0)

static void parse(const std::string &path, rapidjson::Document &doc) {
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) { std::cerr << "cannot open " << path << std::endl; std::exit(100); }
    std::vector<char> buf(65536);
    rapidjson::FileReadStream s(f, buf.data(), buf.size());
    doc.ParseStream(s);
    std::fclose(f);
    T_ASSERT(!doc.HasParseError());
}

static std::vector<real_t> arr(const rapidjson::Document &doc, const char *key) {
    T_ASSERT(doc.HasMember(key) && doc[key].IsArray());
    std::vector<real_t> v(doc[key].Size());
    for (rapidjson::SizeType i = 0; i < doc[key].Size(); i++) v[i] = doc[key][i].GetFloat();
    return v;
}

// Testing.cu:33-76
template <typename T>
static bool host_close(const T *got, const std::vector<real_t> &want, size_t n, double tol = 1e-2) {
    if (want.size() < n) return false;
    for (size_t i = 0; i < n; i++)
        if (std::fabs((double)got[i] - (double)want[i]) > tol) {
            std::cerr << "  mismatch at " << i << ": " << got[i] << " vs " << want[i] << std::endl;
            return false;
        }
    return true;
}
static std::vector<real_t> from_device(const real_t *dev, size_t n) {
    std::vector<real_t> h(n);
    T_ASSERT(cudaMemcpy(h.data(), dev, n * sizeof(real_t), cudaMemcpyDeviceToHost) == cudaSuccess);
    return h;
}
static void to_device(real_t *dev, const std::vector<real_t> &h) {
    T_ASSERT(cudaMemcpy(dev, h.data(), h.size() * sizeof(real_t), cudaMemcpyHostToDevice) == cudaSuccess);
    T_ASSERT(cudaDeviceSynchronize() == cudaSuccess);   // pageable H2D may return before the DMA lands; the library's stream is non-blocking
}
static bool dev_close(const real_t *dev, const std::vector<real_t> &want, size_t off, size_t n, double tol = 1e-2) {
    std::vector<real_t> h = from_device(dev, n);
    std::vector<real_t> w(want.begin() + off, want.begin() + off + n);
    return host_close(h.data(), w, n, tol);
}
// TestSmpcController.cu:28-47
static bool smpc_close(const real_t *dev, const std::vector<real_t> &want) {
    std::vector<real_t> h = from_device(dev, want.size());
    for (size_t i = 0; i < want.size(); i++) {
        const double diff = (double)h[i] - (double)want[i];
        const double measure = std::fabs(h[i]) > 1e2 ? diff / h[i] * 100.0 : diff;
        if (!(std::fabs(measure) < 1e-1)) {
            std::cerr << "  mismatch at " << i << ": " << h[i] << " vs " << want[i] << std::endl;
            return false;
        }
    }
    return true;
}

// ---- loaders (Testing.cu:78-335) ----------------------------------------------------------------------------------------
static int test_loaders(const std::string &cfgPath) {
    SmpcConfiguration cfg(cfgPath);
    rapidjson::Document c;
    parse(cfgPath, c);
    T_ASSERT(cfg.getNX() == (uint_t)arr(c, "nx")[0] && cfg.getNU() == (uint_t)arr(c, "nu")[0]);
    T_ASSERT(cfg.getND() == (uint_t)arr(c, "nd")[0] && cfg.getNV() == (uint_t)arr(c, "nv")[0]);
    T_ASSERT(host_close(cfg.getMatL(), arr(c, "matL"), (size_t)cfg.getNU() * cfg.getNV()));
    T_ASSERT(host_close(cfg.getMatLhat(), arr(c, "matLhat"), (size_t)cfg.getNU() * cfg.getND()));
    T_ASSERT(host_close(cfg.getCostW(), arr(c, "costW"), (size_t)cfg.getNU() * cfg.getNU()));
    T_ASSERT(host_close(cfg.getMatPrcndDiag(), arr(c, "matDiagPrecnd"), arr(c, "matDiagPrecnd").size()));
    T_ASSERT(host_close(cfg.getCurrentX(), arr(c, "currentX"), cfg.getNX()));
    T_ASSERT(host_close(cfg.getPrevU(), arr(c, "prevU"), cfg.getNU()));
    T_ASSERT(host_close(cfg.getPrevDemand(), arr(c, "prevDemand"), cfg.getND()));
    T_ASSERT(cfg.getPenaltyState() == arr(c, "penaltyStateX")[0] && cfg.getPenaltySafety() == arr(c, "penaltySafetyX")[0]);
    T_ASSERT(cfg.getStepSize() == arr(c, "stepSize")[0] && cfg.getMaxIterations() == (uint_t)arr(c, "maxIterations")[0]);
    T_ASSERT(cfg.getWeightEconomical() == 1.0f);
    T_ASSERT(cfg.getPathToNetwork() == c["pathToNetwork"].GetString());

    DwnNetwork net(cfg.getPathToNetwork());
    rapidjson::Document n;
    parse(cfg.getPathToNetwork(), n);
    const size_t nx = net.getNumTanks(), nu = net.getNumControls(), nd = net.getNumDemands(), ne = net.getNumMixNodes();
    T_ASSERT(nx == (size_t)arr(n, "nx")[0] && nu == (size_t)arr(n, "nu")[0] && nd == (size_t)arr(n, "nd")[0] && ne == (size_t)arr(n, "ne")[0]);
    T_ASSERT(host_close(net.getMatA(), arr(n, "matA"), nx * nx));
    T_ASSERT(host_close(net.getMatB(), arr(n, "matB"), nx * nu));
    T_ASSERT(host_close(net.getMatGd(), arr(n, "matGd"), nx * nd));
    T_ASSERT(host_close(net.getMatE(), arr(n, "matE"), ne * nu));
    T_ASSERT(host_close(net.getMatEd(), arr(n, "matEd"), ne * nd));
    T_ASSERT(host_close(net.getXmin(), arr(n, "vecXmin"), nx) && host_close(net.getXmax(), arr(n, "vecXmax"), nx));
    T_ASSERT(host_close(net.getXsafe(), arr(n, "vecXsafe"), nx));
    T_ASSERT(host_close(net.getUmin(), arr(n, "vecUmin"), nu) && host_close(net.getUmax(), arr(n, "vecUmax"), nu));
    T_ASSERT(host_close(net.getAlpha(), arr(n, "costAlpha1"), nu));

    ScenarioTree tree(cfg.getPathToScenarioTree());
    rapidjson::Document t;
    parse(cfg.getPathToScenarioTree(), t);
    const size_t nodes = tree.getNumNodes(), N = tree.getPredHorizon(), K = tree.getNumScenarios();
    T_ASSERT(N == (size_t)arr(t, "N")[0] && K == (size_t)arr(t, "K")[0] && nodes == (size_t)arr(t, "nodes")[0]);
    T_ASSERT(tree.getNumNonleafNodes() == (uint_t)arr(t, "nNonLeafNodes")[0] && tree.getNumChildrenTot() == (uint_t)arr(t, "nChildrenTot")[0]);
    T_ASSERT(host_close(tree.getStageNodes(), arr(t, "stages"), nodes));
    T_ASSERT(host_close(tree.getNodesPerStage(), arr(t, "nodesPerStage"), N));
    T_ASSERT(host_close(tree.getNodesPerStageCumul(), arr(t, "nodesPerStageCumul"), N + 1));
    T_ASSERT(host_close(tree.getLeaveArray(), arr(t, "leaves"), K));
    T_ASSERT(host_close(tree.getChildArray(), arr(t, "children"), tree.getNumChildrenTot()));
    T_ASSERT(host_close(tree.getAncestorArray(), arr(t, "ancestor"), nodes));
    T_ASSERT(host_close(tree.getNumChildren(), arr(t, "nChildren"), tree.getNumNonleafNodes()));
    T_ASSERT(host_close(tree.getNumChildrenCumul(), arr(t, "nChildrenCumul"), nodes));
    T_ASSERT(host_close(tree.getProbArray(), arr(t, "probNode"), nodes));
    T_ASSERT(host_close(tree.getErrorDemandArray(), arr(t, "errorDemandNode"), nodes * nd));
    T_ASSERT(host_close(tree.getErrorPriceArray(), arr(t, "errorPriceNode"), nodes * nu));
    {   // getFinalBranchNode / Stage (ScenarioTree.cu:149-169) against the definition
        std::vector<real_t> nps = arr(t, "nodesPerStage"), cum = arr(t, "nodesPerStageCumul");
        uint_t fbn = 0, fbs = 0;
        for (size_t s = 0; s + 1 < N; s++)
            if (nps[s] == nps[s + 1]) { fbn = (uint_t)cum[s + 1]; fbs = (uint_t)s; break; }
        T_ASSERT(tree.getFinalBranchNode() == fbn && tree.getFinalBranchStage() == fbs);
    }
    bool threw = false;
    try { ScenarioTree missing(cfg.getPathToScenarioTree() + ".does-not-exist"); } catch (const std::logic_error &) { threw = true; }
    T_ASSERT(threw);   // ScenarioTree.cu:40

    Forecaster fc(cfg.getPathToForecaster());
    rapidjson::Document f;
    parse(cfg.getPathToForecaster(), f);
    T_ASSERT(fc.getPredHorizon() == (uint_t)arr(f, "N")[0] && fc.getSimHorizon() == (uint_t)arr(f, "simHorizon")[0]);
    T_ASSERT(fc.getDimDemand() == (uint_t)arr(f, "dimDemand")[0] && fc.getDimPrice() == (uint_t)arr(f, "dimPrices")[0]);
    uint_t slot = 0;
    for (auto it = f.MemberBegin() + 4; it != f.MemberEnd() && it + 1 != f.MemberEnd(); it += 2, slot++) {   // Testing.cu:232-240
        T_ASSERT(fc.predictDemand(slot) == 1 && fc.predictPrices(slot) == 1);
        std::vector<real_t> d(it->value.Size()), p((it + 1)->value.Size());
        for (rapidjson::SizeType i = 0; i < it->value.Size(); i++) d[i] = it->value[i].GetFloat();
        for (rapidjson::SizeType i = 0; i < (it + 1)->value.Size(); i++) p[i] = (it + 1)->value[i].GetFloat();
        T_ASSERT(host_close(fc.getNominalDemand(), d, d.size()) && host_close(fc.getNominalPrices(), p, p.size()));
    }
    T_ASSERT(slot >= 1);
    T_ASSERT(fc.predictDemand(slot + 5) == 0);
    std::cout << "host_tests loaders: ok (" << slot << " forecast slots, " << nodes << " tree nodes)" << std::endl;
    return 0;
}

// ---- Engine + APG steps -------------------------------------------------------------------------------------------------
// TestSmpcController.cuh:80: a subclass that reaches the protected members
class TestSmpcController : public SmpcController {
public:
    using SmpcController::SmpcController;
    int testEngine(const rapidjson::Document &g);
    int testApgSteps(const rapidjson::Document &g);
    int testSurface();
};

static void prepare(TestSmpcController &ctl, const rapidjson::Document &engineGolden) {
    // golden factor matrices are expressed in MATLAB's null-space basis (engineTest.json: matL)
    std::vector<real_t> L = arr(engineGolden, "matL");
    T_ASSERT(rn_set_null_space(ctl.getEngine()->handle(), L.data(), ctl.getSmpcConfiguration()->getMatLhat()) == RN_OK);
    T_ASSERT(ctl.getForecaster()->predictDemand(1) == 1 && ctl.getForecaster()->predictPrices(1) == 1);   // timeInst = 1
    ctl.initialiseSmpcController();
}

int TestSmpcController::testEngine(const rapidjson::Document &g) {
    Engine *e = getEngine();
    DwnNetwork *net = getDwnNetwork();
    ScenarioTree *tree = getScenarioTree();
    const size_t nx = net->getNumTanks(), nu = net->getNumControls(), nv = getSmpcConfiguration()->getNV();
    const size_t nodes = tree->getNumNodes(), N = tree->getPredHorizon(), fbs = tree->getFinalBranchStage();
    T_ASSERT(dev_close(e->getVecUhat(), arr(g, "uHat"), 0, nodes * nu));
    T_ASSERT(dev_close(e->getVecE(), arr(g, "vecE"), 0, nodes * nx));
    T_ASSERT(dev_close(e->getVecBeta(), arr(g, "beta"), 0, nodes * nv));
    T_ASSERT(dev_close(e->getSysMatL(), arr(g, "matL"), 0, nu * nv));
    T_ASSERT(dev_close(e->getPriceAlpha(), arr(g, "costAlpha"), 0, nodes * nu));
    std::vector<real_t> sn = arr(g, "scenarioNodes");
    struct Item { real_t *dev; const char *key; size_t dim; size_t count; };
    const Item items[] = {
        {e->getSysMatF(), "sysF", 2 * nx * nx, N}, {e->getSysMatG(), "sysG", nu * nu, N}, {e->getSysXmin(), "xmin", nx, N},
        {e->getSysXmax(), "xmax", nx, N}, {e->getSysXs(), "xs", nx, N}, {e->getSysUmin(), "umin", nu, N},
        {e->getSysUmax(), "umax", nu, N}, {e->getMatD(), "d", 2 * nx * nv, N}, {e->getMatF(), "f", nu * nv, N},
        {e->getMatPhi(), "Phi", 2 * nx * nv, N}, {e->getMatPsi(), "Psi", nu * nv, N},
        {e->getMatOmega(), "omega", nv * nv, fbs}, {e->getMatTheta(), "Theta", nx * nv, fbs}};
    for (const Item &it : items) {
        std::vector<real_t> want = arr(g, it.key);
        for (size_t k = 0; k < it.count; k++) {   // one representative node per stage (Testing.cu:384-459)
            const size_t node = (size_t)sn[k] - 1;
            if (!dev_close(it.dev + node * it.dim, want, k * it.dim, it.dim)) {
                std::cerr << "engine golden mismatch: " << it.key << " stage " << k << std::endl;
                return 1;
            }
        }
    }
    std::cout << "host_tests engine: ok" << std::endl;
    return 0;
}

int TestSmpcController::testApgSteps(const rapidjson::Document &g) {
    initialiseAlgorithm();
    // testExtrapolation (TestSmpcController.cu:114-168)
    to_device(devVecXi, arr(g, "xi")); to_device(devVecPsi, arr(g, "psi"));
    to_device(devVecUpdateXi, arr(g, "updateXi")); to_device(devVecUpdatePsi, arr(g, "updatePsi"));
    std::vector<real_t> th = arr(g, "theta");
    const real_t lambda = th[1] * (1 / th[0] - 1);
    dualExtrapolationStep(lambda);
    T_ASSERT(smpc_close(devVecAcceleratedXi, arr(g, "acceleXi")) && smpc_close(devVecAcceleratedPsi, arr(g, "accelePsi")));
    T_ASSERT(smpc_close(devVecXi, arr(g, "finalXi")) && smpc_close(devVecPsi, arr(g, "finalPsi")));
    // testSoveStep (:173-216)
    to_device(devVecAcceleratedXi, arr(g, "acceleXi")); to_device(devVecAcceleratedPsi, arr(g, "accelePsi"));
    solveStep();
    T_ASSERT(smpc_close(devVecX, arr(g, "X")) && smpc_close(devVecU, arr(g, "U")));
    // testProximalStep (:221-286)
    to_device(devVecX, arr(g, "X")); to_device(devVecU, arr(g, "U"));
    proximalFunG();
    T_ASSERT(smpc_close(devVecPrimalXi, arr(g, "primalX")) && smpc_close(devVecPrimalPsi, arr(g, "primalU")));
    T_ASSERT(smpc_close(devVecDualXi, arr(g, "dualX")) && smpc_close(devVecDualPsi, arr(g, "dualU")));
    // testFixedPointResidual (:345-398)
    to_device(devVecPrimalXi, arr(g, "primalX")); to_device(devVecPrimalPsi, arr(g, "primalU"));
    to_device(devVecDualXi, arr(g, "dualX")); to_device(devVecDualPsi, arr(g, "dualU"));
    computeFixedPointResidual();
    T_ASSERT(smpc_close(devVecFixedPointResidualXi, arr(g, "fixedPointResidualXi")));
    T_ASSERT(smpc_close(devVecFixedPointResidualPsi, arr(g, "fixedPointResidualPsi")));
    // testDualUpdate (:291-340)
    to_device(devVecFixedPointResidualXi, arr(g, "fixedPointResidualXi"));
    to_device(devVecFixedPointResidualPsi, arr(g, "fixedPointResidualPsi"));
    dualUpdate();
    T_ASSERT(smpc_close(devVecUpdateXi, arr(g, "finalUpdateXi")) && smpc_close(devVecUpdatePsi, arr(g, "finalUpdatePsi")));
    std::cout << "host_tests smpc: ok" << std::endl;
    return 0;
}

// ---- the rest of the Engine / SmpcController surface: per-node pointer tables (Engine.cuh:128-230), the tree on the device
// (Engine.cuh:252-288), the cuBLAS handle, and the stand-alone infeasibility measure against the kernel's own log
int TestSmpcController::testSurface() {
    Engine *e = getEngine();
    ScenarioTree *tree = getScenarioTree();
    const size_t nx = getDwnNetwork()->getNumTanks(), nu = getDwnNetwork()->getNumControls(), nv = getSmpcConfiguration()->getNV();
    const size_t nodes = tree->getNumNodes(), N = tree->getPredHorizon(), K = tree->getNumScenarios(), fb = tree->getFinalBranchNode();
    auto table = [&](real_t **dev, size_t n) {
        std::vector<real_t *> h(n);
        T_ASSERT(cudaMemcpy(h.data(), dev, n * sizeof(real_t *), cudaMemcpyDeviceToHost) == cudaSuccess);
        return h;
    };
    const std::vector<real_t *> phi = table(e->getPtrMatPhi(), nodes), psi = table(e->getPtrMatPsi(), nodes), d = table(e->getPtrMatD(), nodes),
                                f = table(e->getPtrMatF(), nodes), sg = table(e->getPtrMatSigma(), nodes), om = table(e->getPtrMatOmega(), nodes),
                                th = table(e->getPtrMatTheta(), nodes), g = table(e->getPtrMatG(), K), sb = table(e->getPtrSysMatB(), nodes),
                                sl = table(e->getPtrSysMatL(), nodes), slh = table(e->getPtrSysMatLhat(), nodes),
                                sf = table(e->getPtrSysMatF(), nodes), sgg = table(e->getPtrSysMatG(), nodes);
    for (size_t i = 0; i < nodes; i++) {
        T_ASSERT(phi[i] == e->getMatPhi() + i * nv * 2 * nx && psi[i] == e->getMatPsi() + i * nv * nu);
        T_ASSERT(d[i] == e->getMatD() + i * nv * 2 * nx && f[i] == e->getMatF() + i * nv * nu && sg[i] == e->getMatSigma() + i * nv);
        T_ASSERT(sb[i] == e->getSysMatB() && sl[i] == e->getSysMatL() && slh[i] == e->getSysMatLhat());
        T_ASSERT(sf[i] == e->getSysMatF() + i * 2 * nx * nx && sgg[i] == e->getSysMatG() + i * nu * nu);
    }
    for (size_t k = 0; k < K; k++) T_ASSERT(g[k] == e->getMatG() + k * nv * nx);   // one copy per scenario, like the reference
    for (size_t s = 0; s < N; s++)
        for (size_t j = 0; j < (size_t)tree->getNodesPerStage()[s]; j++) {
            const size_t c0 = tree->getNodesPerStageCumul()[s], cur = fb <= c0 ? fb - K + j : c0 + j;   // Engine.cu:210-221
            T_ASSERT(om[c0 + j] == e->getMatOmega() + cur * nv * nv && th[c0 + j] == e->getMatTheta() + cur * nx * nv);
        }
    // the matrices behind the tables are the ones the getters return: Phi of the last node through its table entry
    {
        const std::vector<real_t> a = from_device(phi[nodes - 1], nv * 2 * nx), b = from_device(e->getMatPhi() + (nodes - 1) * nv * 2 * nx, nv * 2 * nx);
        T_ASSERT(std::memcmp(a.data(), b.data(), a.size() * sizeof(real_t)) == 0);
    }
    auto treeu = [&](uint_t *dev, const uint_t *host, size_t n) {
        std::vector<uint_t> h(n);
        T_ASSERT(cudaMemcpy(h.data(), dev, n * sizeof(uint_t), cudaMemcpyDeviceToHost) == cudaSuccess);
        T_ASSERT(std::memcmp(h.data(), host, n * sizeof(uint_t)) == 0);
    };
    treeu(e->getTreeStages(), tree->getStageNodes(), nodes);
    treeu(e->getTreeNodesPerStage(), tree->getNodesPerStage(), N + 1);
    treeu(e->getTreeNodesPerStageCumul(), tree->getNodesPerStageCumul(), N + 2);
    treeu(e->getTreeLeaves(), tree->getLeaveArray(), K);
    treeu(e->getTreeNumChildren(), tree->getNumChildren(), tree->getNumNonleafNodes());
    treeu(e->getTreeAncestor(), tree->getAncestorArray(), nodes);
    treeu(e->getTreeNumChildrenCumul(), tree->getNumChildrenCumul(), nodes);
    {
        const std::vector<real_t> p = from_device(e->getTreeProb(), nodes);
        T_ASSERT(std::memcmp(p.data(), tree->getProbArray(), nodes * sizeof(real_t)) == 0);
        const std::vector<real_t> ed = from_device(e->getTreeErrorDemand(), nodes * getSmpcConfiguration()->getND());
        T_ASSERT(std::memcmp(ed.data(), tree->getErrorDemandArray(), ed.size() * sizeof(real_t)) == 0);
        const std::vector<real_t> ep = from_device(e->getTreeErrorPrices(), nodes * nu);
        T_ASSERT(std::memcmp(ep.data(), tree->getErrorPriceArray(), ep.size() * sizeof(real_t)) == 0);
    }
    T_ASSERT(e->getCublasHandle() != nullptr && e->getCublasHandle() == e->getCublasHandle());
    // updatePrimalInfeasibity (:1480-1496) after the loop == the last entry the kernel logged
    T_ASSERT(algorithmApg() == 1);
    const uint_t iters = getSmpcConfiguration()->getMaxIterations();
    const real_t standalone = updatePrimalInfeasibity(), logged = getPrimalInfeasibility()[iters - 1];
    std::cout << "primal infeasibility: stand-alone " << standalone << ", logged by the kernel " << logged << std::endl;
    T_ASSERT(standalone == logged);
    std::cout << "host_tests surface: ok" << std::endl;
    return 0;
}

// ---- closed loop (main.cu:27-69) -------------------------------------------------------------------------------------------
static int closed_loop(const std::string &cfgPath, int steps, const std::string &outPath) {
    SmpcController *ctl = new SmpcController(cfgPath);
    std::fstream ctrl(outPath + ".control", std::fstream::out);
    std::ofstream out(outPath);
    out << std::setprecision(9);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const uint_t nu = ctl->getSmpcConfiguration()->getNU(), nx = ctl->getSmpcConfiguration()->getNX();
    out << "{\"steps\": [";
    for (int t = 0; t < steps; t++) {
        T_ASSERT(ctl->getForecaster()->predictDemand(t) == 1 && ctl->getForecaster()->predictPrices(t) == 1);
        if (t == 0) ctl->initialiseSmpcController();
        std::vector<real_t> u(nu);
        cudaEventRecord(e0);                                   // tic()  (Utilities.cu:434-447)
        T_ASSERT(ctl->controlAction(u.data()) == 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);         // toc()
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        T_ASSERT(ctl->controlAction(ctrl) == 1);               // the stream variant: clamps u0, fills devControlAction
        ctl->moveForewardInTime();
        out << (t ? ", " : "") << "{\"ms\": " << ms << ", \"u0\": [";
        for (uint_t i = 0; i < nu; i++) out << (i ? ", " : "") << u[i];
        out << "], \"u_applied\": [";
        for (uint_t i = 0; i < nu; i++) out << (i ? ", " : "") << ctl->getSmpcConfiguration()->getPrevU()[i];
        out << "], \"x_next\": [";
        for (uint_t i = 0; i < nx; i++) out << (i ? ", " : "") << ctl->getSmpcConfiguration()->getCurrentX()[i];
        out << "]}";
        std::cout << "time lapsed " << ms << " milliseconds" << std::endl;
    }
    out << "], \"economic_kpi\": " << ctl->getEconomicKpi(steps) << ", \"smooth_kpi\": " << ctl->getSmoothKpi(steps)
        << ", \"safety_kpi\": " << ctl->getSafetyKpi(steps) << ", \"network_kpi\": " << ctl->getNetworkKpi(steps) << "}" << std::endl;
    ctl->releaseObjects();
    delete ctl;
    return 0;
}

// ---- closed-loop lanes: several controllers of ONE GPU side by side (closed-loop Monte-Carlo instances, BASELINE config[3]).
// main.cu:27-63 runs one controller; here `lanes` SmpcController objects, each with its own Engine (= library handle and
// CUDA stream) and its own host thread, run the same loop concurrently.  RAPIDNET_GRID_LIMIT gives each Engine a share of
// the SMs.  All lanes get the same inputs here, so every lane must produce the same controls as lane 0, bit for bit.
static int closed_loop_lanes(const std::string &cfgPath, int steps, int lanes, const std::string &outPath) {
    std::vector<SmpcController *> ctl(lanes, nullptr);
    for (int l = 0; l < lanes; l++) ctl[l] = new SmpcController(cfgPath);
    const uint_t nu = ctl[0]->getSmpcConfiguration()->getNU();
    std::vector<std::vector<real_t>> u0(lanes, std::vector<real_t>((size_t)steps * nu, 0.f));
    std::vector<int> ok(lanes, 1);
    auto run = [&](int l) {
        std::fstream ctrl(outPath + ".control" + std::to_string(l), std::fstream::out);
        for (int t = 0; t < steps; t++) {
            if (ctl[l]->getForecaster()->predictDemand(t) != 1 || ctl[l]->getForecaster()->predictPrices(t) != 1) { ok[l] = 0; return; }
            if (t == 0) ctl[l]->initialiseSmpcController();
            if (ctl[l]->controlAction(u0[l].data() + (size_t)t * nu) != 1) { ok[l] = 0; return; }
            if (ctl[l]->controlAction(ctrl) != 1) { ok[l] = 0; return; }
            ctl[l]->moveForewardInTime();
        }
    };
    run(0);                                                    // warm-up on lane 0 alone (also the sequential timing)
    const auto t0 = std::chrono::steady_clock::now();
    run(0);
    const double one_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const auto t1 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int l = 0; l < lanes; l++) th.emplace_back(run, l);
    for (auto &t : th) t.join();
    const double all_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count();
    int same = 1;
    for (int l = 0; l < lanes; l++) {
        T_ASSERT(ok[l] == 1);
        // lane 0 has moved forward twice before (warm-up + timing): compare the lanes that started from the same state
        if (l > 1 && std::memcmp(u0[l].data(), u0[1].data(), u0[l].size() * sizeof(real_t)) != 0) same = 0;
    }
    std::ofstream out(outPath);
    out << std::setprecision(9) << "{\"lanes\": " << lanes << ", \"steps\": " << steps << ", \"ms_one_lane\": " << one_ms
        << ", \"ms_all_lanes\": " << all_ms << ", \"solves_per_s_one\": " << steps / (one_ms * 1e-3) << ", \"solves_per_s_lanes\": "
        << (double)lanes * steps / (all_ms * 1e-3) << ", \"lanes_identical\": " << same << ", \"u0_lane1\": [";
    const int ref = lanes > 1 ? 1 : 0;
    for (size_t i = 0; i < u0[ref].size(); i++) out << (i ? ", " : "") << u0[ref][i];
    out << "]}" << std::endl;
    for (auto *c : ctl) { c->releaseObjects(); delete c; }
    T_ASSERT(same == 1);
    std::cout << "host_tests lanes: ok (" << lanes << " lanes, " << steps / (one_ms * 1e-3) << " -> " << (double)lanes * steps / (all_ms * 1e-3)
              << " solves/s)" << std::endl;
    return 0;
}

// ---- lazy factor step (SmpcController.cu:579-582): controlAction on a controller that was never initialised -----------------
static int fresh_controller(const std::string &cfgPath) {
    SmpcController lazy(cfgPath), eager(cfgPath);
    const uint_t nu = lazy.getSmpcConfiguration()->getNU();
    std::vector<real_t> a(nu, 0.f), b(nu, 1.f);
    T_ASSERT(lazy.controlAction(a.data()) == 1);            // the reference factors inside solveStep on first use
    eager.initialiseSmpcController();
    T_ASSERT(eager.controlAction(b.data()) == 1);
    T_ASSERT(std::memcmp(a.data(), b.data(), nu * sizeof(real_t)) == 0);
    lazy.releaseObjects(); eager.releaseObjects();
    std::cout << "host_tests fresh: ok" << std::endl;
    return 0;
}

int main(int argc, char **argv) {
    if (argc < 3) { std::cerr << "usage: host_tests loaders|engine|smpc|surface|closedloop|lanes|fresh <controllerConfig.json> ..." << std::endl; return 2; }
    const std::string mode = argv[1], cfg = argv[2];
    if (mode == "loaders") return test_loaders(cfg);
    if (mode == "fresh") return fresh_controller(cfg);
    if (mode == "closedloop") { T_ASSERT(argc >= 5); return closed_loop(cfg, std::atoi(argv[3]), argv[4]); }
    if (mode == "lanes") { T_ASSERT(argc >= 6); return closed_loop_lanes(cfg, std::atoi(argv[3]), std::atoi(argv[4]), argv[5]); }
    T_ASSERT(argc >= 4);
    rapidjson::Document eg;
    parse(argv[3], eg);
    TestSmpcController ctl(cfg);
    prepare(ctl, eg);
    if (mode == "engine") return ctl.testEngine(eg);
    if (mode == "surface") return ctl.testSurface();
    if (mode == "smpc") {
        T_ASSERT(argc >= 5);
        rapidjson::Document sg;
        parse(argv[4], sg);
        return ctl.testApgSteps(sg);
    }
    std::cerr << "unknown mode " << mode << std::endl;
    return 2;
}
