// ---------------------------------------------------------------------------------------------------------
// Derived from openpbso (https://github.com/jhwang7628/openpbso), io.h
//   Copyright (C) 2018 Jui-Hsien Wang <juiwang@alumni.stanford.edu>
// This Source Code Form is subject to the terms of the Mozilla Public License, v. 2.0.  If a copy of the MPL
// was not distributed with this file, You can obtain one at https://mozilla.org/MPL/2.0/.
// The host-side bodies below restate the reference's statements so that a drop-in caller sees bit-identical
// host behaviour (same libstdc++ RNG stream, same state machine); what is new here is the forwarding of the
// hot loops to the B200 C ABI (include/pbso_b200.h).
// ---------------------------------------------------------------------------------------------------------
// openpbso drop-in: the file helpers the synthesis path uses (reference io.h:19-21, io.cpp:18-53) and the pressure-vector
// readers the FFAT fitting step is fed with (reference io.h:24-92).
#ifndef IO_H
#define IO_H
#include <dirent.h>
#include <sys/stat.h>
#include <cassert>
#include <complex>
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <vector>
#include "Eigen/Dense"
namespace Gpu_Wavesolver {
inline bool IsFile(const char* path) { struct stat st; return stat(path, &st) == 0; }
inline std::string Basename(const std::string& path) { return path.substr(path.find_last_of("/") + 1); }
// Appends "<dirname>/<entry>" for every non-dot entry whose full path contains `contains`, in readdir order.
inline void ListDirFiles(const char* dirname, std::vector<std::string>& names, const char* contains = nullptr) {
    DIR* dir = opendir(dirname);
    if (!dir) { perror(""); return; }
    while (struct dirent* ent = readdir(dir)) {
        const std::string f = dirname + std::string("/") + std::string(ent->d_name);
        if (IsFile(f.c_str()) && ent->d_name[0] != '.' && contains && f.find(contains) != std::string::npos)
            names.push_back(f);
    }
    closedir(dir);
}
// Complex vector file: binary = int count (number of reals, 2 per entry) then `count` values of T_i, (re, im)
// pairs; text = one "re im" line per entry (reference io.h:24-65).
template <typename T_i, typename T_o>
void ReadComplexVector(const char* filename, Eigen::Matrix<std::complex<T_o>, Eigen::Dynamic, 1>& p, const bool binary) {
    if (binary) {
        std::ifstream stream(filename, std::ios::binary);
        assert(stream && "file not exist");
        int count = 0;
        stream.read((char*)&count, sizeof(int));
        std::vector<T_i> tmp((size_t)(count > 0 ? count : 0));
        stream.read((char*)tmp.data(), sizeof(T_i) * tmp.size());
        p.resize(count / 2);
        for (int ii = 0; ii < count / 2; ++ii) p(ii) = std::complex<T_o>((T_o)tmp[2 * (size_t)ii], (T_o)tmp[2 * (size_t)ii + 1]);
    } else {
        std::ifstream stream(filename);
        assert(stream && "file not exist");
        std::vector<std::complex<T_o>> rows;
        std::string line;
        while (std::getline(stream, line)) {
            std::istringstream iss(line);
            T_o a = T_o(0), b = T_o(0);
            iss >> a >> b;
            rows.push_back(std::complex<T_o>(a, b));
        }
        p.resize((int)rows.size());
        for (int ii = 0; ii < (int)rows.size(); ++ii) p(ii) = rows[(size_t)ii];
    }
}
// Inverse of ReadComplexVector (reference io.h:67-92): binary writes count = 2 * size then the (re, im) pairs.
template <typename T_i>
void WriteComplexVector(const char* filename, const Eigen::Matrix<std::complex<T_i>, Eigen::Dynamic, 1>& p, const bool binary) {
    if (binary) {
        std::ofstream stream(filename, std::ios::binary);
        const int count = (int)p.size() * 2;
        stream.write((const char*)&count, sizeof(int));
        for (int ii = 0; ii < (int)p.size(); ++ii) {
            const T_i re = p(ii).real(), im = p(ii).imag();
            stream.write((const char*)&re, sizeof(T_i)); stream.write((const char*)&im, sizeof(T_i));
        }
    } else {
        std::ofstream stream(filename);
        stream << std::fixed << std::setprecision(16);
        for (int ii = 0; ii < (int)p.size(); ++ii) stream << p(ii).real() << " " << p(ii).imag() << std::endl;
    }
}
}  // namespace Gpu_Wavesolver
#endif
