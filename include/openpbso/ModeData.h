// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), ModeData.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: ModeData<REAL> (reference ModeData.h:19-148).  Same public members and file format;
// additionally keeps a lazily created device copy of the mode shapes for the impulse projection kernels
// (GetModalForceVertex / GetModalForceFace in modal_force.h).
#ifndef __MODE_DATA_H__
#define __MODE_DATA_H__
#include <cassert>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>
#include <vector>
#include "pbso_check.h"

template <typename REAL>
struct ModeData {
    std::vector<REAL> _omegaSquared;            // eigenvalues
    std::vector<std::vector<REAL>> _modes;      // [nModes][nDOF], mode-major

    int _N_modesAudible = -1;
    REAL _freqThresCache = 22100.;
    REAL _densityCache = -1;

    inline std::vector<REAL>& mode(int modeIndex) { _device.reset(); return _modes.at(modeIndex); }
    inline const std::vector<REAL>& mode(int modeIndex) const { return _modes.at(modeIndex); }
    inline REAL omegaSquared(int modeIndex) const { return _omegaSquared.at(modeIndex); }
    inline int numModes() const { return _omegaSquared.size(); }
    inline int numDOF() const { return (numModes() > 0) ? _modes.at(0).size() : 0; }

    // binary layout: int nDOF, int nModes, REAL w2[nModes], REAL U[nModes][nDOF]  (reference :61-107)
    void read(const char* filename) {
        std::ifstream fin(filename, std::ios::binary);
        assert(fin.good() && "cannot open file for reading modes");
        int nDOF = 0, nModes = 0;
        fin.read((char*)&nDOF, sizeof(int));
        fin.read((char*)&nModes, sizeof(int));
        _omegaSquared.resize(nModes);
        fin.read((char*)_omegaSquared.data(), sizeof(REAL) * nModes);
        _modes.resize(nModes);
        for (auto& m : _modes) { m.resize(nDOF); fin.read((char*)m.data(), sizeof(REAL) * nDOF); }
        _device.reset();
    }
    void write(const char* filename) const {
        std::ofstream fout(filename, std::ios::binary);
        assert(fout.good() && "cannot open file for writing modes");
        const int nModes = _omegaSquared.size(), nDOF = _modes[0].size();
        fout.write((const char*)&nDOF, sizeof(int));
        fout.write((const char*)&nModes, sizeof(int));
        fout.write((const char*)_omegaSquared.data(), sizeof(REAL) * nModes);
        for (const auto& m : _modes) fout.write((const char*)m.data(), sizeof(REAL) * nDOF);
    }
    void printAllFrequency(const REAL& density) const {
        int count = 0;
        for (const REAL& w2 : _omegaSquared) printf("Mode %u: %f Hz\n", count++, sqrt(w2 / density) / (2. * M_PI));
    }
    // Number of leading modes below audibleFreq, with the reference's cache behaviour: only the search
    // loop refreshes the cache, the two early returns do not (reference :120-148).
    int numModesAudible(const REAL& density, const REAL& audibleFreq) {
        if (density == _densityCache && _freqThresCache == audibleFreq && _N_modesAudible >= 0) return _N_modesAudible;
        auto hz = [&](const REAL w2) -> REAL { return sqrt(w2 / density) / (2. * M_PI); };
        if (_omegaSquared.empty() || hz(_omegaSquared.front()) > audibleFreq) return 0;
        if (hz(_omegaSquared.back()) <= audibleFreq) return _omegaSquared.size();
        int n = 0;
        while (n < (int)_omegaSquared.size() && !(hz(_omegaSquared[n]) > audibleFreq)) ++n;
        _N_modesAudible = n; _densityCache = density; _freqThresCache = audibleFreq;
        return n;
    }

    friend std::ostream& operator<<(std::ostream& os, const ModeData& data) {
        os << "------------------------------------------------\nStruct ModeData\n"
           << "------------------------------------------------\n"
           << " number of modes : " << data.numModes() << "\n number of DOF   : " << data.numDOF() << "\n"
           << "------------------------------------------------" << std::flush;
        return os;
    }

    // ---- B200 side: device-resident copy of _modes (uploaded on first use) ----
    pbso_modes* device() const {
        if (!_device) {
            const int M = numModes(), K = numDOF();
            std::vector<double> flat((size_t)M * K);
            for (int m = 0; m < M; ++m)
                for (int d = 0; d < K; ++d) flat[(size_t)m * K + d] = (double)_modes[m][d];
            pbso_modes* h = nullptr;
            pbso_mirror::check(pbso_modes_upload(flat.data(), M, K, &h), "ModeData::device");
            _device = std::shared_ptr<pbso_modes>(h, [](pbso_modes* p) { pbso_modes_destroy(p); });
        }
        return _device.get();
    }
private:
    mutable std::shared_ptr<pbso_modes> _device;
};
#endif  // __MODE_DATA_H__
