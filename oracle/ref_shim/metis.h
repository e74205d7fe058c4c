/* stub: METIS types named by src/Main.hpp:425-431; METIS itself is not used on the solve path */
typedef int idx_t;
typedef float real_t;
