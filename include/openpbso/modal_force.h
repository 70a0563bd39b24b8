// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), tools/real_time_modal_sound.cpp:236-295
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: impulse projection U^T f.  In the reference these two functions live in the GUI tool
// (tools/real_time_modal_sound.cpp:236-295); here they run on the B200 against the device copy of the mode
// shapes (ModeData::device()).  The caller sets force.forceType / force.force afterwards, as the tool does
// from its GUI state (:252-265, :281-294).
#ifndef PBSO_MODAL_FORCE_H
#define PBSO_MODAL_FORCE_H
#include "Eigen/Dense"
#include "ModeData.h"
#include "modal_solver.h"

// force.data[m] = sum_j coords[j] * (vn . U_m[3 vids[j] .. 3 vids[j] + 2]),  m < forceDim
template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
void GetModalForceFace(const int forceDim, const ModeData<T>& modes, const Eigen::Vector3i vids,
                       const Eigen::Vector3d coords, const Eigen::Vector3d& vn, ForceMessage<double, BUF_SIZE>& force) {
    force.data.setZero(forceDim);
    const int v[3] = {vids[0], vids[1], vids[2]};
    const double c[3] = {coords[0], coords[1], coords[2]}, n[3] = {vn[0], vn[1], vn[2]};
    pbso_mirror::check(pbso_modes_project_face(modes.device(), forceDim, v, c, n, force.data.data()), "GetModalForceFace");
}
// force.data[m] = vn . U_m[3 vid .. 3 vid + 2]
template <typename T, int BUF_SIZE = FRAMES_PER_BUFFER>
void GetModalForceVertex(const int forceDim, const ModeData<T>& modes, const int vid, const Eigen::Vector3d& vn,
                         ForceMessage<double, BUF_SIZE>& force) {
    force.data.setZero(forceDim);
    const double n[3] = {vn[0], vn[1], vn[2]};
    pbso_mirror::check(pbso_modes_project_vertex(modes.device(), forceDim, vid, n, force.data.data()), "GetModalForceVertex");
}
#endif
