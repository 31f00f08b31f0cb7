// loaders.cpp -- the four problem-data loaders of the reference (host only, rapidjson), same JSON schema:
//   DwnNetwork          /root/reference/src/DwnNetwork.cu:43-112      (schema DwnNetwork.cuh:23-37)
//   ScenarioTree        /root/reference/src/ScenarioTree.cu:46-121    (schema ScenarioTree.cuh:23-40)
//   Forecaster          /root/reference/src/Forecaster.cu:38-119      (schema Forecaster.cuh:23-30)
//   SmpcConfiguration   /root/reference/src/SmpcConfiguration.cu:44-119 (schema SmpcConfiguration.cuh:24-47)
// All scalars are 1-element arrays, all matrices flat column-major arrays.  Error behaviour as in the reference:
// a missing file prints and exit(100)s (ScenarioTree throws std::logic_error), a missing or non-array key aborts.
// Differences kept on purpose: the whole file is read through a 64 KB buffer (the reference's sizeof(char*) = 8-byte
// buffer is only slow), the per-stage tree arrays are allocated with their real N+1 / N+2 entries, and
// the forecaster copies its time slots at construction instead of moving them out of the DOM on first use.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

// rapidjson is header-only: its inline template members would otherwise be exported from this library and interposed by the
// copies of a program compiled against ANOTHER rapidjson release (the reference vendors v1.1) -- different layouts, a crash
// (the system headers rapidjson pulls in come first, so that only rapidjson's own declarations are hidden)
#include <cassert>
#include <cinttypes>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <iterator>
#include <limits>
#include <memory>
#include <new>
#include <utility>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#pragma GCC visibility push(hidden)
#include "rapidjson/document.h"
#include "rapidjson/filereadstream.h"
#pragma GCC visibility pop

#include "rapidnet_host.hpp"

namespace rapidnet {

namespace {

[[noreturn]] void die(const std::string &msg, int code) {
    std::cerr << msg << std::endl;
    std::exit(code);
}

// parse a whole JSON file; `throws` selects ScenarioTree's behaviour for a missing file
void parse_file(const std::string &path, rapidjson::Document &doc, bool throws = false) {
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) {
        if (throws) throw std::logic_error("Error in opening the file " + path);
        die("Error in opening the file " + path, 100);
    }
    std::vector<char> buffer(65536);
    rapidjson::FileReadStream stream(f, buffer.data(), buffer.size());
    doc.ParseStream(stream);
    std::fclose(f);
    if (doc.HasParseError() || !doc.IsObject()) die("Error in parsing the file " + path, 100);
}

const rapidjson::Value &array_of(const rapidjson::Document &doc, const char *key, const std::string &path) {
    if (!doc.HasMember(key) || !doc[key].IsArray())   // the reference: _ASSERT(a.IsArray()) -> exit(1)
        die("ASSERTION FAILED: \"" + std::string(key) + "\" is not an array in " + path, 1);
    return doc[key];
}

real_t scalar(const rapidjson::Document &doc, const char *key, const std::string &path) {
    const rapidjson::Value &a = array_of(doc, key, path);
    if (a.Size() < 1) die("ASSERTION FAILED: \"" + std::string(key) + "\" is empty in " + path, 1);
    return a[0].GetFloat();
}

void fill(const rapidjson::Value &a, std::vector<real_t> &out) {
    out.resize(a.Size());
    for (rapidjson::SizeType i = 0; i < a.Size(); i++) out[i] = a[i].GetFloat();
}
void fill(const rapidjson::Value &a, std::vector<uint_t> &out) {
    out.resize(a.Size());
    for (rapidjson::SizeType i = 0; i < a.Size(); i++) out[i] = (uint_t)a[i].GetFloat();
}

void expect_size(const std::vector<real_t> &v, size_t n, const char *key, const std::string &path) {
    if (v.size() != n)
        die("ASSERTION FAILED: \"" + std::string(key) + "\" has " + std::to_string(v.size()) + " entries, expected " +
                std::to_string(n) + " in " + path, 1);
}

}  // namespace

// ---- DwnNetwork ------------------------------------------------------------------------------------------------------
DwnNetwork::DwnNetwork(std::string pathToFile) {
    rapidjson::Document doc;
    parse_file(pathToFile, doc);
    nTanks = (uint_t)scalar(doc, "nx", pathToFile);
    nControl = (uint_t)scalar(doc, "nu", pathToFile);
    nDemand = (uint_t)scalar(doc, "nd", pathToFile);
    nMixNodes = (uint_t)scalar(doc, "ne", pathToFile);
    fill(array_of(doc, "matA", pathToFile), matA);   // loaded, never used: A = I is assumed
    fill(array_of(doc, "matB", pathToFile), matB);
    fill(array_of(doc, "matGd", pathToFile), matGd);
    fill(array_of(doc, "matE", pathToFile), matE);
    fill(array_of(doc, "matEd", pathToFile), matEd);
    fill(array_of(doc, "vecXmin", pathToFile), vecXmin);
    fill(array_of(doc, "vecXmax", pathToFile), vecXmax);
    fill(array_of(doc, "vecXsafe", pathToFile), vecXsafe);
    fill(array_of(doc, "vecUmin", pathToFile), vecUmin);
    fill(array_of(doc, "vecUmax", pathToFile), vecUmax);
    fill(array_of(doc, "costAlpha1", pathToFile), vecCostAlpha1);
    expect_size(matB, (size_t)nTanks * nControl, "matB", pathToFile);
    expect_size(matGd, (size_t)nTanks * nDemand, "matGd", pathToFile);
    expect_size(matE, (size_t)nMixNodes * nControl, "matE", pathToFile);
    expect_size(matEd, (size_t)nMixNodes * nDemand, "matEd", pathToFile);
    expect_size(vecXmin, nTanks, "vecXmin", pathToFile);
    expect_size(vecXmax, nTanks, "vecXmax", pathToFile);
    expect_size(vecXsafe, nTanks, "vecXsafe", pathToFile);
    expect_size(vecUmin, nControl, "vecUmin", pathToFile);
    expect_size(vecUmax, nControl, "vecUmax", pathToFile);
    expect_size(vecCostAlpha1, nControl, "costAlpha1", pathToFile);
}

// ---- ScenarioTree ----------------------------------------------------------------------------------------------------
ScenarioTree::ScenarioTree(std::string pathToFile) {
    rapidjson::Document doc;
    parse_file(pathToFile, doc, /*throws=*/true);
    nPredHorizon = (uint_t)scalar(doc, "N", pathToFile);
    nScenario = (uint_t)scalar(doc, "K", pathToFile);
    nNodes = (uint_t)scalar(doc, "nodes", pathToFile);
    nNonLeafNodes = (uint_t)scalar(doc, "nNonLeafNodes", pathToFile);
    nChildrenTot = (uint_t)scalar(doc, "nChildrenTot", pathToFile);
    fill(array_of(doc, "stages", pathToFile), stageArray);
    fill(array_of(doc, "nodesPerStage", pathToFile), nodesPerStage);
    fill(array_of(doc, "nodesPerStageCumul", pathToFile), nodesPerStageCumul);
    fill(array_of(doc, "leaves", pathToFile), leaveArray);
    fill(array_of(doc, "children", pathToFile), childArray);
    fill(array_of(doc, "ancestor", pathToFile), ancestorArray);
    fill(array_of(doc, "nChildren", pathToFile), nChildArray);
    fill(array_of(doc, "nChildrenCumul", pathToFile), nChildCumulArray);
    fill(array_of(doc, "probNode", pathToFile), probNodeArray);
    fill(array_of(doc, "errorDemandNode", pathToFile), errorDemandArray);
    fill(array_of(doc, "errorPriceNode", pathToFile), errorPriceArray);
    // the files carry N+1 / N+2 per-stage entries (trailing 0 / total); tolerate files that stop at N / N+1
    if ((uint_t)nodesPerStage.size() < nPredHorizon + 1) nodesPerStage.resize(nPredHorizon + 1, 0);
    if ((uint_t)nodesPerStageCumul.size() < nPredHorizon + 2) nodesPerStageCumul.resize(nPredHorizon + 2, nNodes);
    if ((uint_t)stageArray.size() != nNodes || (uint_t)ancestorArray.size() != nNodes ||
        (uint_t)probNodeArray.size() != nNodes || (uint_t)leaveArray.size() != nScenario)
        die("ASSERTION FAILED: inconsistent scenario tree arrays in " + pathToFile, 1);
}

uint_t ScenarioTree::getFinalBranchNode() {
    for (uint_t s = 0; s + 1 < nPredHorizon; s++)
        if (nodesPerStage[s] == nodesPerStage[s + 1]) return nodesPerStageCumul[s + 1];
    return 0;
}

uint_t ScenarioTree::getFinalBranchStage() {
    for (uint_t s = 0; s + 1 < nPredHorizon; s++)
        if (nodesPerStage[s] == nodesPerStage[s + 1]) return s;
    return 0;
}

// ---- Forecaster ------------------------------------------------------------------------------------------------------
Forecaster::Forecaster(std::string pathToFile) {
    rapidjson::Document doc;
    parse_file(pathToFile, doc);
    nPredHorizon = (uint_t)scalar(doc, "N", pathToFile);
    simHorizon = (uint_t)scalar(doc, "simHorizon", pathToFile);
    dimDemand = (uint_t)scalar(doc, "dimDemand", pathToFile);
    dimPrices = (uint_t)scalar(doc, "dimPrices", pathToFile);
    nominalDemand.assign((size_t)dimDemand * nPredHorizon, 0.f);
    nominalPrice.assign((size_t)dimPrices * nPredHorizon, 0.f);
    // members 4, 5, 6, ... are the (demand_t, price_t) pairs; their names are irrelevant (Forecaster.cu:93-119)
    uint_t idx = 0;
    for (auto it = doc.MemberBegin(); it != doc.MemberEnd(); ++it, ++idx) {
        if (idx < 4) continue;
        std::vector<real_t> v;
        if (it->value.IsArray()) fill(it->value, v);
        slots.push_back(std::move(v));
    }
}

uint_t Forecaster::predictDemand(uint_t simTime) {
    const size_t k = 2 * (size_t)simTime;
    if (k >= slots.size()) return 0;
    for (size_t i = 0; i < slots[k].size() && i < nominalDemand.size(); i++) nominalDemand[i] = slots[k][i];
    return 1;
}

uint_t Forecaster::predictPrices(uint_t simTime) {
    const size_t k = 2 * (size_t)simTime + 1;
    if (k >= slots.size()) return 0;
    for (size_t i = 0; i < slots[k].size() && i < nominalPrice.size(); i++) nominalPrice[i] = slots[k][i];
    return 1;
}

// ---- SmpcConfiguration -----------------------------------------------------------------------------------------------
SmpcConfiguration::SmpcConfiguration(std::string pathToFile) {
    pathToConfiguration = pathToFile;
    rapidjson::Document doc;
    parse_file(pathToFile, doc);
    NX = (uint_t)scalar(doc, "nx", pathToFile);
    NU = (uint_t)scalar(doc, "nu", pathToFile);
    ND = (uint_t)scalar(doc, "nd", pathToFile);
    NV = (uint_t)scalar(doc, "nv", pathToFile);
    N = (uint_t)scalar(doc, "N", pathToFile);
    fill(array_of(doc, "matL", pathToFile), matL);         // parsed, unused by the Engine (it recomputes L, Lhat)
    fill(array_of(doc, "matLhat", pathToFile), matLhat);
    fill(array_of(doc, "costW", pathToFile), matCostW);
    penaltyStateX = scalar(doc, "penaltyStateX", pathToFile);
    penaltySafetyX = scalar(doc, "penaltySafetyX", pathToFile);
    fill(array_of(doc, "matDiagPrecnd", pathToFile), matDiagPrecnd);
    fill(array_of(doc, "currentX", pathToFile), currentX);
    fill(array_of(doc, "prevU", pathToFile), prevU);
    fill(array_of(doc, "prevDemand", pathToFile), prevDemand);
    stepSize = scalar(doc, "stepSize", pathToFile);
    maxIteration = (uint_t)scalar(doc, "maxIterations", pathToFile);
    auto str = [&](const char *key) -> std::string {
        if (!doc.HasMember(key) || !doc[key].IsString())   // the reference aborts on a missing key (:114-119)
            die("ASSERTION FAILED: \"" + std::string(key) + "\" is not a string in " + pathToFile, 1);
        return doc[key].GetString();
    };
    pathToNetwork = str("pathToNetwork");
    pathToScenarioTree = str("pathToScenarioTree");
    pathToForecaster = str("pathToForecaster");
    algorithmName = str("algorithmName");
    lbfgsBufferSize = (uint_t)scalar(doc, "lbfgsBufferSize", pathToFile);
    expect_size(matCostW, (size_t)NU * NU, "costW", pathToFile);
    expect_size(matDiagPrecnd, (size_t)N * (NU + 2 * NX), "matDiagPrecnd", pathToFile);
    expect_size(currentX, NX, "currentX", pathToFile);
    expect_size(prevU, NU, "prevU", pathToFile);
    expect_size(prevDemand, ND, "prevDemand", pathToFile);
}

void SmpcConfiguration::setCurrentState(real_t *state) { for (uint_t i = 0; i < NX; i++) currentX[i] = state[i]; }
void SmpcConfiguration::setPreviousControl(real_t *control) { for (uint_t i = 0; i < NU; i++) prevU[i] = control[i]; }
void SmpcConfiguration::setpreviousdemand(real_t *demand) { for (uint_t i = 0; i < ND; i++) prevDemand[i] = demand[i]; }

void SmpcConfiguration::setCurrentState() {
    rapidjson::Document doc;
    parse_file(pathToConfiguration, doc);
    const rapidjson::Value &a = array_of(doc, "currentX", pathToConfiguration);
    for (rapidjson::SizeType i = 0; i < a.Size() && i < currentX.size(); i++) currentX[i] = a[i].GetFloat();
}
void SmpcConfiguration::setPreviousControl() {
    rapidjson::Document doc;
    parse_file(pathToConfiguration, doc);
    const rapidjson::Value &a = array_of(doc, "prevU", pathToConfiguration);
    for (rapidjson::SizeType i = 0; i < a.Size() && i < prevU.size(); i++) prevU[i] = a[i].GetFloat();
}
void SmpcConfiguration::setPreviousDemand() {
    // the reference writes the demand it reads into prevU (SmpcConfiguration.cu:290); kept, bounded
    rapidjson::Document doc;
    parse_file(pathToConfiguration, doc);
    const rapidjson::Value &a = array_of(doc, "prevDemand", pathToConfiguration);
    for (rapidjson::SizeType i = 0; i < a.Size() && i < prevU.size(); i++) prevU[i] = a[i].GetFloat();
}

}  // namespace rapidnet
