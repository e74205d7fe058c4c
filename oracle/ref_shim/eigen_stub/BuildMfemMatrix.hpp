/* BuildMfemMatrix.hpp -- TEST INFRASTRUCTURE: src/SMEM_Setup.cpp includes this header, which the reference repository does not contain */
#pragma once
