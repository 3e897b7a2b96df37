! ISO_C_BINDING interface to libdstar_b200.so (include/dstar_b200.h, include/dstar_nscool.h).
!
! This is the module a dStar maintainer adds to NScool (see INTEGRATION.md): with it,
! do_setup_crust_transport (NScool/private/NScool_create_model.f:289) and do_integrate_crust
! (NScool/private/NScool_evolve.f:11) delegate their hot loops to the GPU while everything else in
! NScool_lib stays as it is.  Not compiled in the build image (no Fortran compiler there).
module dstar_b200
   use, intrinsic :: iso_c_binding
   implicit none

   type, bind(C) :: dstar_micro_ctrl
      real(c_double) :: eos_gamma_melt_pt, eos_rsi_melt_pt, eos_nuclide_abundance_threshold
      integer(c_int32_t) :: use_neutron_conductivity, use_superfluid_phonon_conductivity
      integer(c_int32_t) :: use_pcy_for_ee_scattering, use_page_for_eQ_scattering
      real(c_double) :: rad_full_off_lgrho, rad_full_on_lgrho
      integer(c_int32_t) :: use_crust_nu_pair, use_crust_nu_photo, use_crust_nu_plasma
      integer(c_int32_t) :: use_crust_nu_bremsstrahlung, use_crust_nu_pbf
      integer(c_int32_t) :: turn_on_shell_Urca
      real(c_double) :: shell_Urca_luminosity_coeff, lgP_shell_Urca
   end type dstar_micro_ctrl

   type, bind(C) :: dstar_evolve_ctrl
      integer(c_int32_t) :: fix_core_temperature, make_inner_boundary_insulating
      integer(c_int32_t) :: fix_atmosphere_temperature_when_accreting
      integer(c_int32_t) :: turn_on_extra_heating
      integer(c_int32_t) :: suppress_first_step_output, starting_number_for_profile
      integer(c_int32_t) :: maximum_number_of_models
      real(c_double) :: integration_tolerance, min_lg_temperature_integration
      real(c_double) :: max_lg_temperature_integration, maximum_timestep
   end type dstar_evolve_ctrl

   interface
      integer(c_int) function dstar_ctx_create(device, ctx) bind(C, name='dstar_ctx_create')
         import :: c_int, c_ptr
         integer(c_int), value :: device
         type(c_ptr), intent(out) :: ctx
      end function
      subroutine dstar_ctx_destroy(ctx) bind(C, name='dstar_ctx_destroy')
         import :: c_ptr
         type(c_ptr), value :: ctx
      end subroutine
      ! Helm_Table arrays exactly as dStar_eos_def.f:107-133 holds them: pass c_loc of each (imax,jmax) array
      integer(c_int) function dstar_ctx_set_helm(ctx, imax, jmax, arrays) bind(C, name='dstar_ctx_set_helm')
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: imax, jmax
         type(c_ptr), intent(in) :: arrays(21)
      end function
      ! sf_table_type%kF, %f (superfluid_def.f:14-21) after load_sf_table
      integer(c_int) function dstar_ctx_set_gaps(ctx, sf_type, nv, kF_min, kF_max, kF, f4) &
         & bind(C, name='dstar_ctx_set_gaps')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         integer(c_int), value :: sf_type, nv
         real(c_double), value :: kF_min, kF_max
         real(c_double), intent(in) :: kF(*), f4(*)
      end function
      integer(c_int) function dstar_ctx_set_sf_scale(ctx, scale) bind(C, name='dstar_ctx_set_sf_scale')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: ctx
         real(c_double), intent(in) :: scale(3)
      end function
      ! seam A, one-shot host-buffer form.  ionic/ionic_bar: the 11 reals of composition_info_type per zone,
      ! i.e. transfer(s% ionic, [real(dp) ::]) ; all arrays column-major as in NScool_def.f
      integer(c_int) function dstar_micro_tables(ctx, n_star, zone_offset, ncharged, rho, rho_bar, P, P_bar, dm, &
         & ionic, ionic_bar, Yion, Yion_bar, Zion, Aion, Tc_cell, Tc_face, ctrl, tab_lnT, tab_lnCp, tab_lnGamma, &
         & tab_lnEnu, tab_lnK, rho_bar1, status) bind(C, name='dstar_micro_tables')
         import :: c_int, c_int32_t, c_ptr, c_double, dstar_micro_ctrl
         type(c_ptr), value :: ctx
         integer(c_int), value :: n_star, ncharged
         integer(c_int32_t), intent(in) :: zone_offset(*)
         real(c_double), intent(in) :: rho(*), rho_bar(*), P(*), P_bar(*), dm(*), ionic(11,*), ionic_bar(11,*)
         real(c_double), intent(in) :: Yion(ncharged,*), Yion_bar(ncharged,*), Zion(*), Aion(*)
         type(c_ptr), value :: Tc_cell, Tc_face      ! c_loc(Tc(3,nz)) or c_null_ptr
         type(dstar_micro_ctrl), intent(in) :: ctrl
         real(c_double), intent(out) :: tab_lnT(*), tab_lnCp(*), tab_lnGamma(*), tab_lnEnu(*), tab_lnK(*), rho_bar1(*)
         integer(c_int32_t), intent(out) :: status(*)
      end function
      ! seam B: handle-based (dstar_batch_create / upload_structure / upload_tables / upload_evolve / evolve /
      ! download_results); argument lists as in include/dstar_b200.h
      integer(c_int) function dstar_batch_create(ctx, n_star, zone_offset, ncharged, batch) &
         & bind(C, name='dstar_batch_create')
         import :: c_int, c_int32_t, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: n_star, ncharged
         integer(c_int32_t), intent(in) :: zone_offset(*)
         type(c_ptr), intent(out) :: batch
      end function
      subroutine dstar_batch_destroy(batch) bind(C, name='dstar_batch_destroy')
         import :: c_ptr
         type(c_ptr), value :: batch
      end subroutine
      integer(c_int) function dstar_batch_upload_structure(batch, rho, rho_bar, P, P_bar, dm, ionic, ionic_bar, &
         & Yion, Yion_bar, Zion, Aion, Tc_cell, Tc_face) bind(C, name='dstar_batch_upload_structure')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: batch, Tc_cell, Tc_face
         real(c_double), intent(in) :: rho(*), rho_bar(*), P(*), P_bar(*), dm(*), ionic(*), ionic_bar(*)
         real(c_double), intent(in) :: Yion(*), Yion_bar(*), Zion(*), Aion(*)
      end function
      integer(c_int) function dstar_batch_upload_tables(batch, tab_lnT, tab_lnCp, tab_lnGamma, tab_lnEnu, tab_lnK) &
         & bind(C, name='dstar_batch_upload_tables')
         import :: c_int, c_ptr, c_double
         type(c_ptr), value :: batch
         real(c_double), intent(in) :: tab_lnT(*), tab_lnCp(*), tab_lnGamma(*), tab_lnEnu(*), tab_lnK(*)
      end function
      integer(c_int) function dstar_batch_upload_evolve(batch, dm_bar, area, ePhi, e2Phi, ePhi_bar, e2Phi_bar, &
         & n_atm, atm_nv, atm_lgTb, atm_lgTeff, atm_lgflux, atm_index, heat, T_atm, lnT0, n_epoch, &
         & epoch_boundaries, epoch_Mdots, enuc_per_Mdot) bind(C, name='dstar_batch_upload_evolve')
         import :: c_int, c_int32_t, c_ptr, c_double
         type(c_ptr), value :: batch, enuc_per_Mdot
         integer(c_int), value :: n_atm, atm_nv, n_epoch
         real(c_double), intent(in) :: dm_bar(*), area(*), ePhi(*), e2Phi(*), ePhi_bar(*), e2Phi_bar(*)
         real(c_double), intent(in) :: atm_lgTb(*), atm_lgTeff(*), atm_lgflux(*), heat(9,*), T_atm(*), lnT0(*)
         real(c_double), intent(in) :: epoch_boundaries(*), epoch_Mdots(*)
         integer(c_int32_t), intent(in) :: atm_index(*)
      end function
      integer(c_int) function dstar_batch_evolve(batch, ctrl, trace_cap) bind(C, name='dstar_batch_evolve')
         import :: c_int, c_ptr, dstar_evolve_ctrl
         type(c_ptr), value :: batch
         type(dstar_evolve_ctrl), intent(in) :: ctrl
         integer(c_int), value :: trace_cap
      end function
      integer(c_int) function dstar_batch_download_results(batch, t_monitor, Teff_monitor, Qb_monitor, idid, &
         & counters, lnT, Teff, Lsurf, dt, tsec, model) bind(C, name='dstar_batch_download_results')
         import :: c_int, c_ptr
         type(c_ptr), value :: batch, t_monitor, Teff_monitor, Qb_monitor, idid, counters, lnT, Teff, Lsurf, dt, &
            & tsec, model
      end function
      ! history rows and profiles kept on the device (feed do_write_history / do_write_profile)
      integer(c_int) function dstar_batch_download_trace(batch, count, model, tsec, dt, Teff, Lsurf, Lnu, Lnuc, Mdot) &
         & bind(C, name='dstar_batch_download_trace')
         import :: c_int, c_ptr
         type(c_ptr), value :: batch, count, model, tsec, dt, Teff, Lsurf, Lnu, Lnuc, Mdot
      end function
      integer(c_int) function dstar_batch_set_profile_capture(batch, interval, capacity) &
         & bind(C, name='dstar_batch_set_profile_capture')
         import :: c_int, c_ptr
         type(c_ptr), value :: batch
         integer(c_int), value :: interval, capacity
      end function
      integer(c_int) function dstar_batch_download_profiles(batch, count, model, tsec, Mdot, lnT) &
         & bind(C, name='dstar_batch_download_profiles')
         import :: c_int, c_ptr
         type(c_ptr), value :: batch, count, model, tsec, Mdot, lnT
      end function

      ! ---- batch controls added in round 2 (include/dstar_b200.h)
      ! stars with identical coefficient-pass inputs share one table set: owner(i) = 0-based index of the star whose
      ! tables star i uses (owner(i) = i: owns its tables); which = 1 cell tables, 2 face tables
      integer(c_int) function dstar_batch_set_table_sharing(batch, which, owner) &
            & bind(C, name='dstar_batch_set_table_sharing')
         import :: c_int, c_ptr, c_int32_t
         type(c_ptr), value :: batch
         integer(c_int), value :: which
         integer(c_int32_t), intent(in) :: owner(*)
      end function
      ! bc09 envelope tables on the device (dStar_atm_mod.f:46-106): lgTb(nv, n_tab) for n_tab (g, Plight, Pb) triples
      integer(c_int) function dstar_atm_bc09_Teff(ctx, n_tab, grav, Plight, Pb, nv, lgTeff, lgTb, lgflux, ierr, info) &
            & bind(C, name='dstar_atm_bc09_Teff')
         import :: c_int, c_ptr, c_double, c_int32_t
         type(c_ptr), value :: ctx
         integer(c_int), value :: n_tab, nv
         real(c_double), intent(in) :: grav(*), Plight(*), Pb(*), lgTeff(*)
         real(c_double), intent(out) :: lgTb(*), lgflux(*)
         integer(c_int32_t), intent(out) :: ierr(*)
         type(c_ptr), value :: info          ! optional diagnostics (4,nv,n_tab), c_null_ptr to skip
      end function

      ! ---- the sweep driver (include/dstar_nscool.h): NScool handles are the integers alloc_NScool returns
      integer(c_int) function dstar_NScool_create_models(n, ids) bind(C, name='dstar_NScool_create_models')
         import :: c_int, c_int32_t
         integer(c_int), value :: n
         integer(c_int32_t), intent(in) :: ids(*)
      end function
      integer(c_int) function dstar_NScool_evolve_models(n, ids) bind(C, name='dstar_NScool_evolve_models')
         import :: c_int, c_int32_t
         integer(c_int), value :: n
         integer(c_int32_t), intent(in) :: ids(*)
      end function
      ! n stars over n_gpu devices of the node, one host thread per device (replaces n processes looping
      ! NScool_create_model / NScool_evolve_model, NScool_lib.f:40-90); monitors (number_epochs, n) or c_null_ptr
      integer(c_int) function dstar_NScool_run_batch(n_gpu, devices, n, ids, t_monitor, Teff_monitor, Qb_monitor, &
            & star_status) bind(C, name='dstar_NScool_run_batch')
         import :: c_int, c_ptr, c_int32_t
         integer(c_int), value :: n_gpu, n
         type(c_ptr), value :: devices       ! c_loc of an integer(c_int32_t) device list, or c_null_ptr
         integer(c_int32_t), intent(in) :: ids(*)
         type(c_ptr), value :: t_monitor, Teff_monitor, Qb_monitor, star_status
      end function
   end interface
end module dstar_b200
