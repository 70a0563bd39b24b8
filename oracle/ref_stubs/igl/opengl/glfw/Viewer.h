// TEST INFRASTRUCTURE ONLY -- empty stand-in: ffat_solver.h:16 includes the libigl viewer but the
// synthesis path uses nothing from it.
#pragma once
