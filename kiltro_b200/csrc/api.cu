// Library-level entry points: version, error strings, launch counter.
#include "common.cuh"

int64_t g_kb_launches = 0;
int g_kb_pdl = 1;                      // programmatic dependent launch inside the Krylov loops (kb_set_pdl)

int kb_num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
            sms = 148;   // B200
    }
    return sms;
}

extern "C" {

int kb_version(void) { return 100; }

int64_t kb_launch_count(void) { return g_kb_launches; }
void kb_launch_count_reset(void) { g_kb_launches = 0; }
void kb_set_pdl(int on) { g_kb_pdl = on ? 1 : 0; }

const char *kb_error_string(int code) {
    switch (code) {
        case 0: return "success";
        case KB_EINVAL: return "kiltro_b200: invalid argument";
        case KB_ETOOBIG: return "kiltro_b200: count exceeds the int32 index space";
        case KB_EWORKSPACE: return "kiltro_b200: workspace too small";
        case KB_EBREAKDOWN: return "kiltro_b200: Krylov breakdown";
        case KB_ENOTCONV: return "kiltro_b200: Krylov solver did not converge";
        case KB_EPEER: return "kiltro_b200: a peer-memory wait timed out (a rank fell out of the partitioned iteration)";
        case KB_ESTAGNATED: return "kiltro_b200: Krylov solver stagnated above the requested tolerance";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "kiltro_b200: unknown error";
}

}  // extern "C"
