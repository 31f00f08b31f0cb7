// ref_driver.cu -- drives the UNMODIFIED reference (GPUEngineering/RapidNet, compiled from
// /root/reference/src in place) for parity dumps and for the reference arm of bench.py.
// TEST INFRASTRUCTURE ONLY: lives under oracle/, builds into oracle/_ref/, never linked into the product.
//
// It follows the non-test path of the reference's main() (src/main.cu:27-63): construct
// SmpcController(config), predictDemand/Prices(slot), initialiseSmpcController(), then time
// controlAction(real_t*) with CUDA events exactly like tic()/toc() (src/Utilities.cu:428-471).
// A subclass reaches the protected device buffers the same way the reference's own
// TestSmpcController does (src/test/TestSmpcController.cuh:80).
//
// usage: ref_driver <controllerConfig.json> <forecast slot> <warmup> <reps> [dump_dir]
//   prints one line:  REF ms_per_solve=<median> min=<min> iters=<maxIterations> nodes=<nodes> K=<K>
//   dump_dir: raw fp32 files U.bin X.bin V.bin updateXi.bin updatePsi.bin xi.bin psi.bin dualXi.bin dualPsi.bin
//             primalXi.bin primalPsi.bin accelXi.bin accelPsi.bin beta.bin uhat.bin e.bin L.bin Lhat.bin u0.bin pinf.bin
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "cublas_v2.h"
#include "Configuration.h"
#include "SmpcController.cuh"

class RefDriver : public SmpcController {
public:
    explicit RefDriver(std::string cfg) : SmpcController(cfg) {}
    void dump(const std::string &dir) {
        uint_t nx = ptrMyEngine->getDwnNetwork()->getNumTanks();
        uint_t nu = ptrMyEngine->getDwnNetwork()->getNumControls();
        uint_t nd = ptrMyEngine->getDwnNetwork()->getNumDemands();
        uint_t nv = ptrMySmpcConfig->getNV();
        size_t nodes = ptrMyEngine->getScenarioTree()->getNumNodes();
        put(dir, "U", devVecU, nodes * nu); put(dir, "X", devVecX, nodes * nx); put(dir, "V", devVecV, nodes * nv);
        put(dir, "updateXi", devVecUpdateXi, nodes * 2 * nx); put(dir, "updatePsi", devVecUpdatePsi, nodes * nu);
        put(dir, "xi", devVecXi, nodes * 2 * nx); put(dir, "psi", devVecPsi, nodes * nu);
        put(dir, "dualXi", devVecDualXi, nodes * 2 * nx); put(dir, "dualPsi", devVecDualPsi, nodes * nu);
        put(dir, "primalXi", devVecPrimalXi, nodes * 2 * nx); put(dir, "primalPsi", devVecPrimalPsi, nodes * nu);
        put(dir, "accelXi", devVecAcceleratedXi, nodes * 2 * nx); put(dir, "accelPsi", devVecAcceleratedPsi, nodes * nu);
        put(dir, "beta", ptrMyEngine->getVecBeta(), nodes * nv); put(dir, "uhat", ptrMyEngine->getVecUhat(), nodes * nu);
        put(dir, "e", ptrMyEngine->getVecE(), nodes * nx);
        put(dir, "L", ptrMyEngine->getSysMatL(), (size_t)nu * nv); put(dir, "Lhat", ptrMyEngine->getSysMatLhat(), (size_t)nu * nd);
        FILE *f = fopen((dir + "/pinf.bin").c_str(), "wb");
        if (f) { fwrite(vecPrimalInfs, sizeof(real_t), ptrMySmpcConfig->getMaxIterations(), f); fclose(f); }
    }
private:
    static void put(const std::string &dir, const char *name, real_t *dev, size_t count) {
        std::vector<real_t> h(count);
        _CUDA(cudaMemcpy(h.data(), dev, count * sizeof(real_t), cudaMemcpyDeviceToHost));
        FILE *f = fopen((dir + "/" + name + ".bin").c_str(), "wb");
        if (!f) { fprintf(stderr, "ref_driver: cannot write %s/%s.bin\n", dir.c_str(), name); exit(3); }
        fwrite(h.data(), sizeof(real_t), count, f);
        fclose(f);
    }
};

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: %s config.json slot warmup reps [dump_dir]\n", argv[0]); return 2; }
    const std::string cfg = argv[1];
    const int slot = atoi(argv[2]), warmup = atoi(argv[3]), reps = atoi(argv[4]);
    const std::string dump = argc > 5 ? argv[5] : "";
    RefDriver *c = new RefDriver(cfg);
    c->getForecaster()->predictDemand(slot);
    c->getForecaster()->predictPrices(slot);
    c->initialiseSmpcController();
    const uint_t nu = c->getDwnNetwork()->getNumControls();
    std::vector<real_t> u0(nu);
    cudaEvent_t e0, e1;
    _CUDA(cudaEventCreate(&e0)); _CUDA(cudaEventCreate(&e1));
    std::vector<float> ms;
    for (int r = 0; r < warmup + reps; r++) {
        _CUDA(cudaDeviceSynchronize());
        _CUDA(cudaEventRecord(e0, 0));
        c->controlAction(u0.data());          // src/SmpcController.cu:1607-1625
        _CUDA(cudaEventRecord(e1, 0));
        _CUDA(cudaEventSynchronize(e1));
        float t = 0.f;
        _CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (r >= warmup) ms.push_back(t);
        fprintf(stderr, "ref_driver: solve %d: %.3f ms\n", r, t);
    }
    if (!dump.empty()) {
        c->dump(dump);
        FILE *f = fopen((dump + "/u0.bin").c_str(), "wb");
        if (f) { fwrite(u0.data(), sizeof(real_t), nu, f); fclose(f); }
    }
    std::sort(ms.begin(), ms.end());
    const float med = ms.empty() ? 0.f : ms[ms.size() / 2], mn = ms.empty() ? 0.f : ms[0];
    printf("REF ms_per_solve=%.6f min=%.6f iters=%d nodes=%d K=%d reps=%d\n", med, mn,
           (int)c->getSmpcConfiguration()->getMaxIterations(), (int)c->getScenarioTree()->getNumNodes(),
           (int)c->getScenarioTree()->getNumScenarios(), reps);
    return 0;
}
