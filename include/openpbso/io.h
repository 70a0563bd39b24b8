// openpbso drop-in: the file helpers the synthesis path uses (reference io.h:19-21, io.cpp:18-53).
#ifndef IO_H
#define IO_H
#include <dirent.h>
#include <sys/stat.h>
#include <cstdio>
#include <string>
#include <vector>
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
}  // namespace Gpu_Wavesolver
#endif
