// hostnuma.hpp — host-side placement helpers: which NUMA node a CUDA device hangs off, binding the calling thread to
// that node's cores, and pinned host buffers whose pages live on the node of the device that will copy them.
//
// Why: the end-to-end path of the batched configs is PCIe-bound (SCALE_r01: 52 GB/s H2D at one GPU). With several GPUs in
// one box every copy is fed from host DRAM; pages that sit on the other socket cross the inter-socket link first. The
// reference fans channels out over goroutines inside one call (constant.go:223-241); here the fan-out is over devices,
// one bound worker thread per device (capi.cu), and the caller's buffers can be allocated shard-by-shard on the right node.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

namespace gar {

// NUMA node of CUDA device `device` from sysfs (/sys/bus/pci/devices/<bdf>/numa_node); -1 when unknown / single node.
int device_numa_node(int device);
// CPUs of a node (/sys/devices/system/node/node<N>/cpulist); empty when unknown.
std::vector<int> node_cpus(int node);
// Restrict the calling thread to the CPUs of `device`'s node. Returns the node (>= 0) or -1 when nothing was changed.
int bind_thread_to_device(int device);
// Human-readable summary ("dev 0 -> node 0 (cpus 0-55)") for logs and the bench line.
std::string describe_placement(int device);

}  // namespace gar
