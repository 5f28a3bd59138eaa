// no MPI in the test build
