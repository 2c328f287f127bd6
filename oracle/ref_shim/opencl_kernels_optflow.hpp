// generated OpenCL kernel table in a real build; the shim build has no OpenCL
