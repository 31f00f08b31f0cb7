// shim_main.cu -- the in-scope part of the reference's own main() (/root/reference/src/main.cu:12-19, TESTING = 1) for the
// build of oracle/build_shim_tests.sh: the reference's UNCHANGED test classes, compiled against rapidnet-b200's class surface.
// Out of scope here, as in DESIGN.md: testSmpcFbeController / testSmpcNamaController (main.cu:22-23: the FBE / NAMA solvers).
// Run from a directory whose ../test/testDataFiles holds the reference's test JSON files (tests/test_host_cpp.py writes them
// from tests/golden/toy.npz).  TEST INFRASTRUCTURE ONLY.
#include "SmpcController.cuh"
#include "test/Testing.cuh"

int main(void) {
    Testing *myTesting = new Testing();
    _ASSERT(myTesting->testNetwork());
    _ASSERT(myTesting->testScenarioTree());
    _ASSERT(myTesting->testForecaster());
    _ASSERT(myTesting->testControllerConfig());
    _ASSERT(myTesting->testEngineTesting());
    _ASSERT(myTesting->testSmpcController());
    cout << "ref_shim_tests: the reference's loader, Engine and APG tests pass against rapidnet-b200" << endl;
    return 0;
}
