// TEST INFRASTRUCTURE ONLY -- empty stand-in for ffat_solver.h:17 (unused on the synthesis path).
#pragma once
