!> iso_c_binding interface to libc2ray_b200.so (include/c2ray_b200.h) and the replacement body of
!! subroutine evolve3D (code/files_for_3D/evolve.F90:78) that forwards to it.
!!
!! STATUS: delivered as source only.  The build image has no Fortran compiler (SURVEY F1), so this file has not been
!! compiled; the same ABI is exercised from C (tests/test_abi_cpu.py struct-layout check) and from ctypes
!! (c2-ray3dm1d_helium_b200/capi.py), which use the identical memory layout (column-major grids, 1-based srcpos).
!!
!! How a maintainer uses it (see INTEGRATION.md):
!!   1. add this file to the Makefile's object list and link with -lc2ray_b200 -lcudart
!!   2. in evolve.F90 replace the body of evolve3D by `call evolve3D_b200(time,dt,restart)`
!!   3. call c2ray_b200_setup() once after rad_ini(), setup_cool(), evolve_ini() and the source list are available,
!!      and c2ray_b200_new_sources() whenever sourceprops changes the source list
module c2ray_b200

  use, intrinsic :: iso_c_binding
  implicit none
  private

  public :: c2ray_b200_setup, c2ray_b200_new_sources, evolve3D_b200, c2ray_b200_finish

  integer(c_int), parameter :: C2RAY_NUMFREQBND = 47, C2RAY_MAX_ITER_HIST = 512

  type, bind(C) :: c2ray_params
     integer(c_int32_t) :: isothermal, cosmological, subboxsize, max_subbox
     real(c_double) :: temper_val, H0, Omega0
     real(c_float) :: clumping
     integer(c_int32_t) :: max_slots, deterministic
  end type c2ray_params

  type, bind(C) :: c2ray_sed_tables
     type(c_ptr) :: photo_thick, photo_thin, heat_thick, heat_thin
     integer(c_int32_t) :: freqbnd_lower, freqbnd_upper
     real(c_double) :: S_star
  end type c2ray_sed_tables

  type, bind(C) :: c2ray_stats
     integer(c_int32_t) :: niter, conv_flag, conv_criterion, nit_max
     integer(c_int64_t) :: sum_nbox_all, rt_updates, chem_cells, nit_total
     real(c_double) :: photon_loss_all, ms_sweep, ms_chem, ms_allreduce, ms_total
     real(c_double) :: sums_before(5), sums_after(5)
     real(c_double) :: totrec, totcollisions, recomions, total_ion, totalsrc, photcons
     integer(c_int32_t) :: conv_hist(C2RAY_MAX_ITER_HIST)
  end type c2ray_stats

  interface
     function c2ray_b200_init(params, mesh, device, ctx) bind(C, name="c2ray_b200_init") result(rc)
       import :: c_int, c_int32_t, c_ptr, c2ray_params
       type(c2ray_params), intent(in) :: params
       integer(c_int32_t), intent(in) :: mesh(3)
       integer(c_int32_t), value :: device
       type(c_ptr), intent(out) :: ctx
       integer(c_int) :: rc
     end function c2ray_b200_init
     function c2ray_b200_destroy(ctx) bind(C, name="c2ray_b200_destroy") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: rc
     end function c2ray_b200_destroy
     function c2ray_b200_set_cooling_tables(ctx, logT, logL) bind(C, name="c2ray_b200_set_cooling_tables") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(in) :: logT(801), logL(801,5)
       integer(c_int) :: rc
     end function c2ray_b200_set_cooling_tables
     function c2ray_b200_upload_tables(ctx, sed, tables) bind(C, name="c2ray_b200_upload_tables") result(rc)
       import :: c_int, c_int32_t, c_ptr, c2ray_sed_tables
       type(c_ptr), value :: ctx
       integer(c_int32_t), value :: sed
       type(c2ray_sed_tables), intent(in) :: tables
       integer(c_int) :: rc
     end function c2ray_b200_upload_tables
     function c2ray_b200_set_sources(ctx, NumSrc, srcpos, NormFlux, NormFluxPL, NormFluxQPL) &
          bind(C, name="c2ray_b200_set_sources") result(rc)
       import :: c_int, c_int32_t, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int32_t), value :: NumSrc
       integer(c_int32_t), intent(in) :: srcpos(3,*)
       real(c_double), intent(in) :: NormFlux(*)
       type(c_ptr), value :: NormFluxPL, NormFluxQPL   ! c_loc(array(1)) or c_null_ptr
       integer(c_int) :: rc
     end function c2ray_b200_set_sources
     function c2ray_b200_set_geometry(ctx, dr, vol, zred) bind(C, name="c2ray_b200_set_geometry") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(in) :: dr(3)
       real(c_double), value :: vol, zred
       integer(c_int) :: rc
     end function c2ray_b200_set_geometry
     function c2ray_b200_evolve3d_host(ctx, time, dt, restart, ndens, xh, xhe, temperature_grid, stats) &
          bind(C, name="c2ray_b200_evolve3d_host") result(rc)
       import :: c_int, c_int32_t, c_ptr, c_double, c_float, c2ray_stats
       type(c_ptr), value :: ctx
       real(c_double), value :: time, dt
       integer(c_int32_t), value :: restart
       real(c_double), intent(in) :: ndens(*)
       real(c_double), intent(inout) :: xh(*), xhe(*)
       type(c_ptr), value :: temperature_grid          ! c_loc(temperature_grid) or c_null_ptr when isothermal
       type(c2ray_stats), intent(out) :: stats
       integer(c_int) :: rc
     end function c2ray_b200_evolve3d_host
     function c2ray_b200_get_rates(ctx, phih, phihe, phiheat) bind(C, name="c2ray_b200_get_rates") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: phih(*), phihe(*), phiheat(*)
       integer(c_int) :: rc
     end function c2ray_b200_get_rates
     function c2ray_b200_get_work_state(ctx, xh_av, xhe_av, xh_intermed, xhe_intermed) &
          bind(C, name="c2ray_b200_get_work_state") result(rc)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: xh_av(*), xhe_av(*), xh_intermed(*), xhe_intermed(*)
       integer(c_int) :: rc
     end function c2ray_b200_get_work_state
     function c2ray_b200_comm_unique_id(id) bind(C, name="c2ray_b200_comm_unique_id") result(rc)
       import :: c_int, c_int8_t
       integer(c_int8_t), intent(out) :: id(128)
       integer(c_int) :: rc
     end function c2ray_b200_comm_unique_id
     function c2ray_b200_comm_init(ctx, id, rank, npr) bind(C, name="c2ray_b200_comm_init") result(rc)
       import :: c_int, c_int8_t, c_int32_t, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int8_t), intent(in) :: id(128)
       integer(c_int32_t), value :: rank, npr
       integer(c_int) :: rc
     end function c2ray_b200_comm_init
     function c2ray_b200_set_dump(ctx, dump_dir, interval_s) bind(C, name="c2ray_b200_set_dump") result(rc)
       import :: c_int, c_ptr, c_char, c_double
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: dump_dir(*)   ! trim(adjustl(dump_dir))//c_null_char
       real(c_double), value :: interval_s
       integer(c_int) :: rc
     end function c2ray_b200_set_dump
     function c2ray_b200_set_clumping_grid(ctx, clumping_grid) bind(C, name="c2ray_b200_set_clumping_grid") result(rc)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr), value :: clumping_grid                  ! c_loc(clumping_grid) (default real) or c_null_ptr
       integer(c_int) :: rc
     end function c2ray_b200_set_clumping_grid
     function c2ray_b200_set_LLS(ctx, type_of_LLS, coldensh_LLS, LLS_grid) bind(C, name="c2ray_b200_set_LLS") result(rc)
       import :: c_int, c_int32_t, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int32_t), value :: type_of_LLS
       real(c_double), value :: coldensh_LLS
       type(c_ptr), value :: LLS_grid                       ! c_loc(LLS_grid) (default real) or c_null_ptr
       integer(c_int) :: rc
     end function c2ray_b200_set_LLS
     function c2ray_b200_write_stream2(ctx, results_dir, zred_now) bind(C, name="c2ray_b200_write_stream2") result(rc)
       import :: c_int, c_ptr, c_char, c_double
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: results_dir(*)
       real(c_double), value :: zred_now
       integer(c_int) :: rc
     end function c2ray_b200_write_stream2
     function c2ray_b200_write_stream3(ctx, results_dir, zred_now) bind(C, name="c2ray_b200_write_stream3") result(rc)
       import :: c_int, c_ptr, c_char, c_double
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: results_dir(*)
       real(c_double), value :: zred_now
       integer(c_int) :: rc
     end function c2ray_b200_write_stream3
     function c2ray_b200_last_error() bind(C, name="c2ray_b200_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function c2ray_b200_last_error
  end interface

  type(c_ptr), save :: ctx = c_null_ptr

contains

  !> Address of a host array that the reference declares without TARGET (temperature_grid, NormFluxQPL, ...): c_loc needs
  !! a TARGET, and associating the actual argument with a TARGET dummy supplies one without touching the reference's
  !! declarations.  The library only reads/writes through the pointer during the call it is passed to.
  function loc_sp(a) result(p)
    real(c_float), target, intent(in) :: a(*)
    type(c_ptr) :: p
    p = c_loc(a)
  end function loc_sp

  function loc_dp(a) result(p)
    real(c_double), target, intent(in) :: a(*)
    type(c_ptr) :: p
    p = c_loc(a)
  end function loc_dp

  !> Stop with the library's message: the reference has no error returns, it logs and stops (SURVEY 8b).
  subroutine check(rc, where)
    use file_admin, only: logf
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    if (rc /= 0) then
       call c_f_pointer(c2ray_b200_last_error(), msg, [256])
       write(logf,*) "c2ray_b200 error ", rc, " in ", where, ": ", msg
       stop "c2ray_b200 error"
    endif
  end subroutine check

  !> Once per run, after rad_ini / setup_cool / evolve_ini: creates the device context, uploads the radiation tables the
  !! host's own radiation_tables module built, the cooling curves, and attaches the NCCL communicator that replaces
  !! the MPI_ALLREDUCE calls of evolve.F90:505-548.
  subroutine c2ray_b200_setup()
    use precision, only: dp
    use my_mpi, only: rank, npr
#ifdef MPI
    use my_mpi, only: MPI_COMM_NEW
#endif
    use sizes, only: mesh
    use c2ray_parameters, only: subboxsize, max_subbox, cosmological
    use material, only: isothermal, temper_val, clumping
    use cosmology_parameters, only: H0, Omega0
    use radiation_tables, only: bb_photo_thick_table, bb_photo_thin_table, bb_heat_thick_table, bb_heat_thin_table, &
         bb_FreqBnd_LowerLimit, bb_FreqBnd_UpperLimit
    use radiation_sed_parameters, only: S_star
    use file_admin, only: dump_dir
#ifdef PL
    use radiation_tables, only: pl_photo_thick_table, pl_photo_thin_table, pl_heat_thick_table, &
         pl_heat_thin_table, pl_FreqBnd_LowerLimit, pl_FreqBnd_UpperLimit
    use radiation_sed_parameters, only: pl_S_star
#endif
#ifdef QUASARS
    use radiation_tables, only: qpl_photo_thick_table, qpl_photo_thin_table, qpl_heat_thick_table, &
         qpl_heat_thin_table, qpl_FreqBnd_LowerLimit, qpl_FreqBnd_UpperLimit
    use radiation_sed_parameters, only: qpl_S_star
#endif
#ifdef MPI
    include 'mpif.h'
    integer :: ierr
#endif
    type(c2ray_params) :: par
    type(c2ray_sed_tables) :: tab
    integer(c_int8_t) :: uid(128)
    real(c_double) :: logT(801), logL(801,5)

    par%isothermal = merge(1, 0, isothermal)
    par%cosmological = merge(1, 0, cosmological)
    par%subboxsize = subboxsize
    par%max_subbox = max_subbox
    par%temper_val = real(temper_val, c_double)
    par%H0 = H0
    par%Omega0 = Omega0
    par%clumping = clumping
    par%max_slots = 0
    par%deterministic = 0
    call check(c2ray_b200_init(par, int(mesh, c_int32_t), -1_c_int32_t, ctx), "init")

    ! cooling curves exactly as cooling_h.f90:83-149 reads them (log10 values; the library applies 10**x)
    call read_cooling_logs(logT, logL)
    call check(c2ray_b200_set_cooling_tables(ctx, logT, logL), "set_cooling_tables")

    tab%photo_thick = c_loc(bb_photo_thick_table)
    tab%photo_thin = c_loc(bb_photo_thin_table)
    if (isothermal) then
       tab%heat_thick = c_null_ptr
       tab%heat_thin = c_null_ptr
    else
       tab%heat_thick = c_loc(bb_heat_thick_table)
       tab%heat_thin = c_loc(bb_heat_thin_table)
    endif
    tab%freqbnd_lower = bb_FreqBnd_LowerLimit
    tab%freqbnd_upper = bb_FreqBnd_UpperLimit
    tab%S_star = S_star
    call check(c2ray_b200_upload_tables(ctx, 0_c_int32_t, tab), "upload_tables(B)")
#ifdef PL
    ! the power-law SED "P" (radiation_photoionrates.f90:214-220 reads these tables with NormFluxPL): sed index 1
    tab%photo_thick = c_loc(pl_photo_thick_table)
    tab%photo_thin = c_loc(pl_photo_thin_table)
    if (.not.isothermal) then
       tab%heat_thick = c_loc(pl_heat_thick_table)
       tab%heat_thin = c_loc(pl_heat_thin_table)
    endif
    tab%freqbnd_lower = pl_FreqBnd_LowerLimit
    tab%freqbnd_upper = pl_FreqBnd_UpperLimit
    tab%S_star = pl_S_star
    call check(c2ray_b200_upload_tables(ctx, 1_c_int32_t, tab), "upload_tables(P)")
#endif
#ifdef QUASARS
    tab%photo_thick = c_loc(qpl_photo_thick_table)
    tab%photo_thin = c_loc(qpl_photo_thin_table)
    if (.not.isothermal) then
       tab%heat_thick = c_loc(qpl_heat_thick_table)
       tab%heat_thin = c_loc(qpl_heat_thin_table)
    endif
    tab%freqbnd_lower = qpl_FreqBnd_LowerLimit
    tab%freqbnd_upper = qpl_FreqBnd_UpperLimit
    tab%S_star = qpl_S_star
    call check(c2ray_b200_upload_tables(ctx, 2_c_int32_t, tab), "upload_tables(Q)")
#endif

    ! iteration dumps every 15 minutes into dump_dir, as evolve.F90:199-213 does; evolve3D's restart flag then finds
    ! iterdump1.bin / iterdump2.bin / iterdump.bin there (evolve.F90:279 start_from_dump)
    call check(c2ray_b200_set_dump(ctx, trim(adjustl(dump_dir))//c_null_char, 15.0_c_double*60.0_c_double), "set_dump")

#ifdef MPI
    if (npr > 1) then
       if (rank == 0) call check(c2ray_b200_comm_unique_id(uid), "comm_unique_id")
       call MPI_BCAST(uid, 128, MPI_BYTE, 0, MPI_COMM_NEW, ierr)
       call check(c2ray_b200_comm_init(ctx, uid, int(rank, c_int32_t), int(npr, c_int32_t)), "comm_init")
    endif
#endif
  end subroutine c2ray_b200_setup

  !> The five curves of cooling_h.f90:83-149 as log10 values (the module itself keeps only 10**x, privately).
  subroutine read_cooling_logs(logT, logL)
    real(c_double), intent(out) :: logT(801), logL(801,5)
    character(len=40), parameter :: files(5) = [character(len=40) :: "../tables/H0-cool.tab", &
         "../tables/H1-cool-B.tab", "../tables/He0-cool_new.tab", "../tables/He1-cool_new_nocollion.tab", &
         "../tables/He2-cool.tab"]
    integer :: n, i, element, ion, nchck
    real(c_double) :: t
    do n = 1, 5
       open(unit=22, file=trim(files(n)), status='old')
       read(22,*) element, ion, nchck
       do i = 1, 801
          read(22,*) t, logL(i,n)
          if (n == 1) logT(i) = t
       enddo
       close(22)
    enddo
  end subroutine read_cooling_logs

  !> After sourceprops has (re)built the source list.
  subroutine c2ray_b200_new_sources()
    use sourceprops, only: NumSrc, srcpos, NormFlux
#ifdef PL
    use sourceprops, only: NormFluxPL
#endif
#ifdef QUASARS
    use sourceprops, only: NormFluxQPL
#endif
    type(c_ptr) :: pp, pq
    pp = c_null_ptr
    pq = c_null_ptr
#ifdef PL
    if (NumSrc > 0) pp = loc_dp(NormFluxPL)       ! sourceprops_test.F90:136-162
#endif
#ifdef QUASARS
    if (NumSrc > 0) pq = loc_dp(NormFluxQPL)
#endif
    ! NormFlux is dimensioned (0:NumSrc) in the reference: pass element 1 onwards
    call check(c2ray_b200_set_sources(ctx, int(NumSrc, c_int32_t), int(srcpos, c_int32_t), NormFlux(1:NumSrc), &
         pp, pq), "set_sources")
  end subroutine c2ray_b200_new_sources

  !> Drop-in body of evolve3D(time,dt,restart), code/files_for_3D/evolve.F90:78-229.
  subroutine evolve3D_b200(time, dt, restart)
    use precision, only: dp
    use file_admin, only: logf
    use my_mpi, only: rank
    use grid, only: dr, vol
    use cosmology, only: zred
    use material, only: ndens, xh, xhe, temperature_grid, isothermal
    ! position-dependent clumping and Lyman-limit systems: module variables of material (public by default:
    ! mat_ini_test.F90:37,45, mat_ini_cubep3m.F90:43,53), refreshed by set_clumping / set_LLS once per redshift slice
    use material, only: clumping_grid, LLS_grid, coldensh_LLS
    use c2ray_parameters, only: type_of_clumping, use_LLS, type_of_LLS
    use evolve_data, only: phih_grid, phihe_grid, phiheat, xh_av, xhe_av, xh_intermed, xhe_intermed, photon_loss_all
    use evolve_source, only: sum_nbox_all
    use sizes, only: mesh
    use photonstatistics, only: photon_loss, totrec, totcollisions, recomions, dh0, dhe0, dhe2, total_ion, &
         report_photonstatistics, update_grandtotal_photonstatistics
    real(kind=dp), intent(in) :: time, dt
    integer, intent(in) :: restart
    type(c2ray_stats) :: st
    type(c_ptr) :: pt

    call check(c2ray_b200_set_geometry(ctx, dr, vol, zred), "set_geometry")   ! dr, vol, zred change every step
    ! evolve_point.F90:484 clumping_point / :177-180 LLS_point: the per-cell values live on the device
    if (type_of_clumping == 5) then
       call check(c2ray_b200_set_clumping_grid(ctx, loc_sp(clumping_grid)), "set_clumping_grid")
    endif
    if (use_LLS) then
       if (type_of_LLS == 2) then
          call check(c2ray_b200_set_LLS(ctx, 2_c_int32_t, 0.0_c_double, loc_sp(LLS_grid)), "set_LLS")
       else
          call check(c2ray_b200_set_LLS(ctx, 1_c_int32_t, coldensh_LLS, c_null_ptr), "set_LLS")
       endif
    endif
    pt = c_null_ptr
    if (.not.isothermal) pt = loc_sp(temperature_grid)
    call check(c2ray_b200_evolve3d_host(ctx, time, dt, int(restart, c_int32_t), ndens, xh, xhe, pt, st), "evolve3d")
    ! what output.F90:354,364 and the final photon statistics (evolve.F90:225) read afterwards
    call check(c2ray_b200_get_rates(ctx, phih_grid, phihe_grid, phiheat), "get_rates")
    call check(c2ray_b200_get_work_state(ctx, xh_av, xhe_av, xh_intermed, xhe_intermed), "get_work_state")
    photon_loss_all(:) = 0.0_dp
    photon_loss_all(1) = st%photon_loss_all
    sum_nbox_all = int(st%sum_nbox_all)
    ! Photon statistics of the step (evolve.F90:225-227).  calculate_photon_statistics cannot be called on the host: its
    ! total_rates reads the recombination coefficients the last do_chemistry call left in the cgsconstants module
    ! (photonstatistics.f90:180-194), and that call now happened on the device.  The library evaluates the same sums
    ! with the same coefficients; the module's public variables are filled from them and the host's own reporting runs.
    photon_loss(:) = photon_loss_all(:)/(real(mesh(1))*real(mesh(2))*real(mesh(3)))   ! evolve.F90:457
    totrec = st%totrec
    totcollisions = st%totcollisions
    recomions = st%recomions
    dh0 = st%sums_before(1) - st%sums_after(1)      ! photonstatistics.f90:251-260 total_ionizations
    dhe0 = st%sums_before(3) - st%sums_after(3)
    dhe2 = st%sums_after(5) - st%sums_before(5)
    total_ion = st%total_ion
    call report_photonstatistics (dt)
    call update_grandtotal_photonstatistics (dt)
    if (rank == 0) then
       write(logf,*) "Multiple sources convergence reached after ", st%niter, " iterations"
       write(logf,*) "Test 1 values: ", st%conv_flag, st%conv_criterion
    endif
  end subroutine evolve3D_b200

  subroutine c2ray_b200_finish()
    integer(c_int) :: rc
    if (c_associated(ctx)) rc = c2ray_b200_destroy(ctx)
    ctx = c_null_ptr
  end subroutine c2ray_b200_finish

end module c2ray_b200
