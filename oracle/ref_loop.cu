// ref_loop.cu -- the closed loop of the UNMODIFIED reference (GPUEngineering/RapidNet, compiled from /root/reference/src in
// place): the non-test branch of its main() (src/main.cu:27-63) as a function of the step count, with the controls, states
// and KPIs written out for comparison.  TEST INFRASTRUCTURE ONLY: lives under oracle/, builds into oracle/_ref/.
//
//   predictDemand/Prices(t) -> [t == 0: initialiseSmpcController()] -> controlAction(fstream&) -> moveForewardInTime()
//   then getEconomicKpi / getSmoothKpi / getSafetyKpi / getNetworkKpi (src/SmpcController.cu:1778-1859)
//
// usage: ref_loop <controllerConfig.json> <steps> <out.json>
//   out.json: {"steps": [{"u_applied": [...], "x_next": [...]}, ...], "economic_kpi": .., "smooth_kpi": .., "safety_kpi": ..,
//              "network_kpi": ..}   (u_applied / x_next = what moveForewardInTime leaves in the configuration: prevU, currentX)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <string>

#include "cublas_v2.h"
#include "Configuration.h"
#include "SmpcController.cuh"

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s config.json steps out.json\n", argv[0]); return 2; }
    const std::string cfg = argv[1], outPath = argv[3];
    const int steps = atoi(argv[2]);
    SmpcController *c = new SmpcController(cfg);
    std::fstream ctrl((outPath + ".control").c_str(), std::fstream::out);
    std::ofstream out(outPath.c_str());
    out << std::setprecision(9);
    const uint_t nu = c->getSmpcConfiguration()->getNU(), nx = c->getSmpcConfiguration()->getNX();
    out << "{\"steps\": [";
    for (int t = 0; t < steps; t++) {
        c->getForecaster()->predictDemand(t);
        c->getForecaster()->predictPrices(t);
        if (t == 0) c->initialiseSmpcController();
        c->controlAction(ctrl);                 // src/SmpcController.cu:1633-1667 (clamps u0, fills devControlAction)
        c->moveForewardInTime();                // :1679-1717
        out << (t ? ", " : "") << "{\"u_applied\": [";
        for (uint_t i = 0; i < nu; i++) out << (i ? ", " : "") << c->getSmpcConfiguration()->getPrevU()[i];
        out << "], \"x_next\": [";
        for (uint_t i = 0; i < nx; i++) out << (i ? ", " : "") << c->getSmpcConfiguration()->getCurrentX()[i];
        out << "]}";
    }
    out << "], \"economic_kpi\": " << c->getEconomicKpi(steps) << ", \"smooth_kpi\": " << c->getSmoothKpi(steps)
        << ", \"safety_kpi\": " << c->getSafetyKpi(steps) << ", \"network_kpi\": " << c->getNetworkKpi(steps) << "}" << std::endl;
    out.close();
    printf("REF_LOOP steps=%d ok\n", steps);
    return 0;
}
