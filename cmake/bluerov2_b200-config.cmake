# bluerov2_b200-config.cmake -- locates a tree staged by `python -m bluerov2_b200.build --stage DIR`.
#
#   find_package(bluerov2_b200 CONFIG REQUIRED PATHS DIR/cmake)
#
# sets the three variables bluerov2_dobmpc/CMakeLists.txt:39-41 defines by hand, so that the rest of that file
# (include_directories :71-78, link_directories :87, target_link_libraries :94-97,108-111,128-132,158-162) works as is:
#   acados_include  = DIR/acados/include        (was "~/acados/include")
#   acados_lib      = DIR/acados/lib            (was "~/acados/lib")
#   bluerov2_model  = DIR/c_generated_code      (was ${PROJECT_SOURCE_DIR}/scripts/c_generated_code)
# and, for new consumers, the imported targets
#   bluerov2_b200::solver   libacados_ocp_solver_bluerov2.so + all include directories
#   bluerov2_b200::acados   the libacados.so link shim (carries a dependency on the solver library)
get_filename_component(_br2_root "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(acados_include "${_br2_root}/acados/include")
set(acados_lib "${_br2_root}/acados/lib")
set(bluerov2_model "${_br2_root}/c_generated_code")
foreach(_f "${bluerov2_model}/libacados_ocp_solver_bluerov2.so" "${bluerov2_model}/acados_solver_bluerov2.h"
           "${acados_lib}/libacados.so" "${acados_include}/acados_c/ocp_nlp_interface.h")
  if(NOT EXISTS "${_f}")
    set(bluerov2_b200_FOUND FALSE)
    set(bluerov2_b200_NOT_FOUND_MESSAGE "bluerov2_b200: ${_f} is missing (stage the tree with `python -m bluerov2_b200.build --stage DIR`)")
    return()
  endif()
endforeach()
if(NOT TARGET bluerov2_b200::solver)
  add_library(bluerov2_b200::solver SHARED IMPORTED)
  set_target_properties(bluerov2_b200::solver PROPERTIES
    IMPORTED_LOCATION "${bluerov2_model}/libacados_ocp_solver_bluerov2.so"
    IMPORTED_SONAME "libacados_ocp_solver_bluerov2.so"
    INTERFACE_INCLUDE_DIRECTORIES "${bluerov2_model};${acados_include};${acados_include}/blasfeo/include")
  add_library(bluerov2_b200::acados SHARED IMPORTED)
  set_target_properties(bluerov2_b200::acados PROPERTIES
    IMPORTED_LOCATION "${acados_lib}/libacados.so"
    IMPORTED_SONAME "libacados.so"
    INTERFACE_LINK_LIBRARIES bluerov2_b200::solver)
endif()
set(bluerov2_b200_FOUND TRUE)
