/* TEST INFRASTRUCTURE ONLY (oracle build shim). */
#ifndef OPENMM_EXPORT
#define OPENMM_EXPORT
#endif
