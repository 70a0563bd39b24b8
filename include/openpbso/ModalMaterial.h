// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), ModalMaterial.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: ModalMaterial<REAL> (reference ModalMaterial.h:19-56).  Host-only carrier; nothing here
// needs the device.
#ifndef MODAL_MATERIAL_H
#define MODAL_MATERIAL_H
#include <cmath>
#include <fstream>
#include <sstream>
#include <string>
#include "config.h"
template <typename REAL>
struct ModalMaterial {
    std::string name;
    REAL alpha;
    REAL beta;
    REAL density;
    REAL poissonRatio;
    REAL youngsModulus;
    REAL inverseDensity;   // declared by the reference, never set there either (ModalMaterial.h:29)

    // Rayleigh damping ratio and damped frequency (DyRT eq. 10, 12; reference :30-33)
    inline REAL xi(const REAL& omega_i) const { return 0.5 * (alpha / omega_i + beta * omega_i); }
    inline REAL omega_di(const REAL& omega_i) const { return omega_i * sqrt(1.0 - pow(xi(omega_i), 2)); }

    // Text format: leading '#' lines, then "density E nu alpha beta" (reference :35-55).
    // Returns a new'd object, or nullptr when the file cannot be opened.
    static ModalMaterial* Read(const char* filename) {
        std::ifstream in(filename);
        if (!in) return nullptr;
        ModalMaterial* mat = new ModalMaterial();
        mat->name = filename;
        std::string line;
        while (std::getline(in, line))
            if (line[0] != '#') break;
        std::istringstream fields(line);
        fields >> mat->density >> mat->youngsModulus >> mat->poissonRatio >> mat->alpha >> mat->beta;
        return mat;
    }
};
#endif
