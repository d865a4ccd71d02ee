// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// A do-nothing stand-in for the Khronos OpenCL C++ bindings, just big enough that the reference's
// algorithms/MSV_HMM.cpp compiles *unmodified* in an image that has no OpenCL headers or ICD.  Only the
// reference's sequential path (MSV_HMM::run_on_sequence, MSV_HMM.cpp:74-113) is ever executed from the object
// built with this header; every OpenCL call below is inert and reports failure.  The names and numeric values of
// the constants are the ones published in the OpenCL 1.2 specification (they are needed because the reference
// switches over them in get_error_string, MSV_HMM.cpp:121-195).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

using cl_int = int32_t;
using cl_uint = uint32_t;
using cl_float = float;
using cl_bool = cl_uint;
using cl_mem_flags = uint64_t;
using cl_device_type = uint64_t;

enum : cl_int {
    CL_SUCCESS = 0,
    CL_DEVICE_NOT_FOUND = -1,
    CL_DEVICE_NOT_AVAILABLE = -2,
    CL_COMPILER_NOT_AVAILABLE = -3,
    CL_MEM_OBJECT_ALLOCATION_FAILURE = -4,
    CL_OUT_OF_RESOURCES = -5,
    CL_OUT_OF_HOST_MEMORY = -6,
    CL_PROFILING_INFO_NOT_AVAILABLE = -7,
    CL_MEM_COPY_OVERLAP = -8,
    CL_IMAGE_FORMAT_MISMATCH = -9,
    CL_IMAGE_FORMAT_NOT_SUPPORTED = -10,
    CL_BUILD_PROGRAM_FAILURE = -11,
    CL_MAP_FAILURE = -12,
    CL_MISALIGNED_SUB_BUFFER_OFFSET = -13,
    CL_EXEC_STATUS_ERROR_FOR_EVENTS_IN_WAIT_LIST = -14,
    CL_COMPILE_PROGRAM_FAILURE = -15,
    CL_LINKER_NOT_AVAILABLE = -16,
    CL_LINK_PROGRAM_FAILURE = -17,
    CL_DEVICE_PARTITION_FAILED = -18,
    CL_KERNEL_ARG_INFO_NOT_AVAILABLE = -19,
    CL_INVALID_VALUE = -30,
    CL_INVALID_DEVICE_TYPE = -31,
    CL_INVALID_PLATFORM = -32,
    CL_INVALID_DEVICE = -33,
    CL_INVALID_CONTEXT = -34,
    CL_INVALID_QUEUE_PROPERTIES = -35,
    CL_INVALID_COMMAND_QUEUE = -36,
    CL_INVALID_HOST_PTR = -37,
    CL_INVALID_MEM_OBJECT = -38,
    CL_INVALID_IMAGE_FORMAT_DESCRIPTOR = -39,
    CL_INVALID_IMAGE_SIZE = -40,
    CL_INVALID_SAMPLER = -41,
    CL_INVALID_BINARY = -42,
    CL_INVALID_BUILD_OPTIONS = -43,
    CL_INVALID_PROGRAM = -44,
    CL_INVALID_PROGRAM_EXECUTABLE = -45,
    CL_INVALID_KERNEL_NAME = -46,
    CL_INVALID_KERNEL_DEFINITION = -47,
    CL_INVALID_KERNEL = -48,
    CL_INVALID_ARG_INDEX = -49,
    CL_INVALID_ARG_VALUE = -50,
    CL_INVALID_ARG_SIZE = -51,
    CL_INVALID_KERNEL_ARGS = -52,
    CL_INVALID_WORK_DIMENSION = -53,
    CL_INVALID_WORK_GROUP_SIZE = -54,
    CL_INVALID_WORK_ITEM_SIZE = -55,
    CL_INVALID_GLOBAL_OFFSET = -56,
    CL_INVALID_EVENT_WAIT_LIST = -57,
    CL_INVALID_EVENT = -58,
    CL_INVALID_OPERATION = -59,
    CL_INVALID_GL_OBJECT = -60,
    CL_INVALID_BUFFER_SIZE = -61,
    CL_INVALID_MIP_LEVEL = -62,
    CL_INVALID_GLOBAL_WORK_SIZE = -63,
    CL_INVALID_PROPERTY = -64,
    CL_INVALID_IMAGE_DESCRIPTOR = -65,
    CL_INVALID_COMPILER_OPTIONS = -66,
    CL_INVALID_LINKER_OPTIONS = -67,
    CL_INVALID_DEVICE_PARTITION_COUNT = -68,
    CL_INVALID_GL_SHAREGROUP_REFERENCE_KHR = -1000,
    CL_PLATFORM_NOT_FOUND_KHR = -1001,
    CL_INVALID_D3D10_DEVICE_KHR = -1002,
    CL_INVALID_D3D10_RESOURCE_KHR = -1003,
    CL_D3D10_RESOURCE_ALREADY_ACQUIRED_KHR = -1004,
    CL_D3D10_RESOURCE_NOT_ACQUIRED_KHR = -1005,
};

constexpr cl_bool CL_TRUE = 1;
constexpr cl_uint CL_PLATFORM_NAME = 0x0902;
constexpr cl_uint CL_DEVICE_NAME = 0x102B;
constexpr cl_uint CL_PROGRAM_BUILD_LOG = 0x1183;
constexpr cl_device_type CL_DEVICE_TYPE_DEFAULT = 1u << 0;
constexpr cl_mem_flags CL_MEM_READ_WRITE = 1u << 0;
constexpr cl_mem_flags CL_MEM_READ_ONLY = 1u << 2;
constexpr cl_mem_flags CL_MEM_USE_HOST_PTR = 1u << 3;
constexpr cl_mem_flags CL_MEM_HOST_READ_ONLY = 1u << 8;
constexpr cl_mem_flags CL_MEM_HOST_NO_ACCESS = 1u << 9;

namespace cl {

struct Device {
    template <cl_uint What> std::string getInfo() const { return "no-opencl-stub"; }
};

struct Platform {
    // One fake platform so that `platforms[0]` (MSV_HMM.cpp:213) is a valid element.
    static cl_int get(std::vector<Platform>* out) {
        out->emplace_back();
        return CL_SUCCESS;
    }
    cl_int getInfo(cl_uint, std::string* out) const {
        *out = "stub";
        return CL_SUCCESS;
    }
    cl_int getDevices(cl_device_type, std::vector<Device>*) const { return CL_DEVICE_NOT_FOUND; }
};

struct Context {
    Context() = default;
    Context(const std::vector<Device>&, const void*, const void*, const void*, cl_int* err) {
        if (err) *err = CL_DEVICE_NOT_AVAILABLE;
    }
};

struct Buffer {
    Buffer() = default;
    Buffer(const Context&, cl_mem_flags, std::size_t, void*, cl_int* err) {
        if (err) *err = CL_INVALID_CONTEXT;
    }
};

struct NDRange {
    NDRange() = default;
    explicit NDRange(std::size_t) {}
};
static const NDRange NullRange;

struct Program {
    Program() = default;
    Program(const Context&, const std::string&, bool, cl_int* err) {
        if (err) *err = CL_INVALID_CONTEXT;
    }
    cl_int build(const char*) { return CL_INVALID_PROGRAM; }
    template <cl_uint What> std::vector<std::pair<Device, std::string>> getBuildInfo() const { return {}; }
};

struct Kernel {
    Kernel() = default;
    Kernel(const Program&, const char*, cl_int* err) {
        if (err) *err = CL_INVALID_PROGRAM_EXECUTABLE;
    }
    template <class T> cl_int setArg(cl_uint, const T&) { return CL_INVALID_KERNEL; }
};

struct CommandQueue {
    CommandQueue() = default;
    CommandQueue(const Context&, cl_uint, cl_int* err) {
        if (err) *err = CL_INVALID_CONTEXT;
    }
    cl_int enqueueNDRangeKernel(const Kernel&, const NDRange&, const NDRange&, const NDRange&) {
        return CL_INVALID_COMMAND_QUEUE;
    }
    cl_int enqueueReadBuffer(const Buffer&, cl_bool, std::size_t, std::size_t, void*) {
        return CL_INVALID_COMMAND_QUEUE;
    }
};

} // namespace cl
