// oracle/shim/thrust/device_vector.h — intentionally empty (bvh.h:5 includes it, uses nothing from it)
