// hostnuma.cpp — see hostnuma.hpp. Linux sysfs + sched_setaffinity only (no libnuma in the image).
#include "hostnuma.hpp"

#include <cuda_runtime.h>
#include <sched.h>

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace gar {

static bool read_line(const std::string& path, std::string& out) {
    std::ifstream f(path);
    if (!f) return false;
    std::getline(f, out);
    return true;
}

int device_numa_node(int device) {
    char bdf[32] = {0};
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    std::string id(bdf);
    std::transform(id.begin(), id.end(), id.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    std::string line;
    if (!read_line("/sys/bus/pci/devices/" + id + "/numa_node", line)) return -1;
    return std::atoi(line.c_str());  // the kernel reports -1 on single-node machines
}

std::vector<int> node_cpus(int node) {  // "0-55,112-167"
    std::vector<int> cpus;
    std::string line;
    if (node < 0 || !read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", line)) return cpus;
    std::stringstream ss(line);
    std::string part;
    while (std::getline(ss, part, ',')) {
        int a = 0, b = 0;
        const int k = std::sscanf(part.c_str(), "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1)
            for (int c = a; c <= b; ++c) cpus.push_back(c);
    }
    return cpus;
}

int bind_thread_to_device(int device) {
    const int node = device_numa_node(device);
    const std::vector<int> cpus = node_cpus(node);
    if (cpus.empty()) return -1;
    // keep only CPUs the process is allowed to use (cgroup / taskset limits)
    cpu_set_t allowed, want;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return -1;
    CPU_ZERO(&want);
    int n = 0;
    for (int c : cpus)
        if (c < CPU_SETSIZE && CPU_ISSET(c, &allowed)) {
            CPU_SET(c, &want);
            ++n;
        }
    if (n == 0) return -1;
    if (sched_setaffinity(0, sizeof(want), &want) != 0) return -1;
    return node;
}

std::string describe_placement(int device) {
    const int node = device_numa_node(device);
    std::string line;
    std::string s = "dev " + std::to_string(device) + " -> node " + std::to_string(node);
    if (node >= 0 && read_line("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist", line)) s += " (cpus " + line + ")";
    return s;
}

}  // namespace gar
