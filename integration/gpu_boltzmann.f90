! gpu_boltzmann.f90 -- the Fortran side of the seam between CAMB's cmbmain and libcambb200.so.
!
! New file for fortran/ (add to the SOURCEFILES list of Makefile_main after equations.f90 / before cmbmain.f90, link
! with -L$(CAMBB200)/camb_b200 -lcambb200 -lcudart).  Together with integration/cmbmain.patch it redirects the four
! OpenMP loops of cmbmain.f90 (:205-209 DoSourcek, :1147-1163 TransferOut, :287-291 SourceToTransfers, :322-323
! CalcLimberScalCls + CalcScalCls) to the C-ABI of include/cambb200.h.  Everything above cmbmain -- CAMB_GetResults,
! camb_python.f90, the Python package, the ini driver -- is untouched: the host arrays the library fills are the
! reference's own allocatables (ThisSources%LinearSrc, State%MT%TransferData, ThisCT%Delta_p_l_k, iCl_scalar, iCl_Array).
!
! The bind(C) types are generated from the C headers (integration/gen_seam.py -> gpu_boltzmann_types.inc) and their
! field offsets are asserted in C (integration/abi_offsets_test.c, run by tests/test_integration_files.py); this image
! has no Fortran compiler, so the module itself has not been compiled here -- it is written against the reference
! names of results.f90 / equations.f90 / cmbmain.f90 (v2.0.x), cited inline.
module GpuBoltzmann
    use, intrinsic :: iso_c_binding
    use precision
    use results
    use config
    use constants
    implicit none
    private

    integer, parameter :: CB_MAX_NU = 5, CB_MAX_NQ = 16, CB_MAX_WINDOWS = 8
    include 'gpu_boltzmann_types.inc'

    logical, public :: use_gpu_boltzmann = .true.      ! ini key / CAMBparams flag in a full integration
    logical, public :: build_bessel_on_device = .true.  ! skip GenerateBessels on the host: cb_build_bessel (bessels.f90:140-228 on the device)
    logical, public :: thermo_tables_on_device = .true. ! leave results.f90:2430-2494 (derivative tables, Horner packing) to the device
    type(c_ptr), save :: ctx = c_null_ptr              ! one context per process (cmbmain is not re-entrant, cmbmain.f90:7-8)
    type(c_ptr), save :: dm = c_null_ptr               ! device copy of the current CAMBdata
    ! per-call arrays the cb_model / cb_modes structs point to (kept alive until the calls return)
    real(c_double), allocatable, target, save :: tight(:), nuapp(:, :), pows(:), limb_pows(:, :), nu_notmassless(:, :)
    type(cb_range), allocatable, target, save :: tau_regions(:), bess_regions(:)
    type(cb_window), allocatable, target, save :: windows(:)

    public :: GpuBoltzmann_Init, GpuBoltzmann_Free, GpuCalcSources, GpuTransferOut, GpuSourceToTransfers, GpuCalcScalCls, &
        GpuCalcTensCls

    interface
        integer(c_int) function cb_create(device, ctx) bind(C, name='cb_create')
            import; integer(c_int), value :: device; type(c_ptr) :: ctx
        end function
        subroutine cb_destroy(ctx) bind(C, name='cb_destroy')
            import; type(c_ptr), value :: ctx
        end subroutine
        type(c_ptr) function cb_last_error(ctx) bind(C, name='cb_last_error')
            import; type(c_ptr), value :: ctx
        end function
        integer(c_int) function cb_set_bessel(ctx, b) bind(C, name='cb_set_bessel')
            import; type(c_ptr), value :: ctx; type(cb_bessel), intent(in) :: b
        end function
        integer(c_int) function cb_build_bessel(ctx, b, horner_out) bind(C, name='cb_build_bessel')   ! GenerateBessels on the device
            import; type(c_ptr), value :: ctx; type(cb_bessel), intent(in) :: b; type(c_ptr), value :: horner_out
        end function
        integer(c_int) function cb_model_create(ctx, m, dm) bind(C, name='cb_model_create')
            import; type(c_ptr), value :: ctx; type(cb_model), intent(in) :: m; type(c_ptr) :: dm
        end function
        subroutine cb_model_destroy(ctx, dm) bind(C, name='cb_model_destroy')
            import; type(c_ptr), value :: ctx, dm
        end subroutine
        integer(c_int) function cb_evolve(ctx, n_models, dm, modes, LinearSrc, TransferData, stats) bind(C, name='cb_evolve')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: n_models
            type(c_ptr), intent(in) :: dm(*); type(cb_modes), intent(in) :: modes(*)
            type(c_ptr), intent(in) :: LinearSrc(*), TransferData(*), stats(*)
        end function
        integer(c_int) function cb_set_sources(ctx, dm, n_k, q, ScaledSrc) bind(C, name='cb_set_sources')
            import; type(c_ptr), value :: ctx, dm; integer(c_int), value :: n_k; type(c_ptr), value :: q, ScaledSrc
        end function
        integer(c_int) function cb_make_nonlinear_sources(ctx, dm, n_k, n_z, nonlin_ratio, Min_kh_nonlinear) &
                bind(C, name='cb_make_nonlinear_sources')                                            ! MakeNonlinearSources on the device
            import; type(c_ptr), value :: ctx, dm; integer(c_int), value :: n_k, n_z; type(c_ptr), value :: nonlin_ratio
            real(c_double), value :: Min_kh_nonlinear
        end function
        integer(c_int) function cb_comm_unique_id(ctx, id128) bind(C, name='cb_comm_unique_id')     ! MPI builds: one cosmology over
            import; type(c_ptr), value :: ctx; character(kind=c_char), intent(out) :: id128(128)     ! the GPUs of a box (k-sharding)
        end function
        integer(c_int) function cb_comm_init(ctx, nranks, rank, id128) bind(C, name='cb_comm_init')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: nranks, rank; character(kind=c_char), intent(in) :: id128(128)
        end function
        integer(c_int) function cb_gather_sources(ctx, dm, n_k, owner) bind(C, name='cb_gather_sources')
            import; type(c_ptr), value :: ctx, dm; integer(c_int), value :: n_k; integer(c_int), intent(in) :: owner(*)
        end function
        integer(c_int) function cb_allreduce_cl(ctx, dm, iCl) bind(C, name='cb_allreduce_cl')
            import; type(c_ptr), value :: ctx, dm, iCl
        end function
        integer(c_int) function cb_project(ctx, n_models, dm, n_q, q_int, dq, Delta) bind(C, name='cb_project')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: n_models
            type(c_ptr), intent(in) :: dm(*); integer(c_int), intent(in) :: n_q(*); type(c_ptr), intent(in) :: q_int(*), dq(*), Delta(*)
        end function
        integer(c_int) function cb_cl(ctx, n_models, dm, tensor, pows, iCl) bind(C, name='cb_cl')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: n_models, tensor
            type(c_ptr), intent(in) :: dm(*), pows(*), iCl(*)
        end function
        integer(c_int) function cb_cl_array(ctx, n_models, dm, pows, iCl_Array) bind(C, name='cb_cl_array')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: n_models
            type(c_ptr), intent(in) :: dm(*), pows(*), iCl_Array(*)
        end function
        integer(c_int) function cb_cl_windows(ctx, n_models, dm, pows, limber_pows, iCl, iCl_Array) bind(C, name='cb_cl_windows')
            import; type(c_ptr), value :: ctx; integer(c_int), value :: n_models
            type(c_ptr), intent(in) :: dm(*), pows(*), limber_pows(*), iCl(*), iCl_Array(*)
        end function
    end interface

contains

    subroutine GpuBoltzmann_Init(device)
        integer, intent(in), optional :: device
        integer(c_int) :: err, dev
        if (c_associated(ctx)) return
        dev = 0; if (present(device)) dev = device
        err = cb_create(dev, ctx)
        if (err /= 0) call GlobalError('GPU Boltzmann path: no usable CUDA device (there is no CPU fallback)', error_unsupported_params)
    end subroutine

    subroutine GpuBoltzmann_Free
        if (c_associated(ctx)) then
            if (c_associated(dm)) call cb_model_destroy(ctx, dm)
            call cb_destroy(ctx)
            ctx = c_null_ptr; dm = c_null_ptr
        end if
    end subroutine

    subroutine Check(err, where)
        ! device-side errors carry CAMB's own ids (config.f90:62-72): 4 = error_evolution, 5 = error_unsupported_params
        integer(c_int), intent(in) :: err
        character(LEN=*), intent(in) :: where
        character(kind=c_char), pointer :: msg(:)
        character(LEN=256) :: text
        integer :: i
        if (err == 0) return
        text = ''
        call c_f_pointer(cb_last_error(ctx), msg, [256])
        do i = 1, 256
            if (msg(i) == c_null_char) exit
            text(i:i) = msg(i)
        end do
        call GlobalError(where//': '//trim(text), int(err))
    end subroutine

    ! ---- the two per-k host look-ups of `initial`, hoisted (cb_modes.tight_switchoff / nu_massive_approx_time) ----

    function TightSwitchoffTimeFor(State, k) result(t)
        ! equations.f90:1935-1950 with ep0 = 1e-2 (:186), epsw = 100/tau0 (:539) and second_order_tightcoupling = .true. (:52)
        type(CAMBdata), intent(in), target :: State
        real(dl), intent(in) :: k
        real(dl) :: t, ep, epsw
        real(dl), parameter :: ep0 = 1.0d-2
        epsw = 100/State%tau0
        if (k > epsw) then
            if (k > epsw*5) then
                ep = ep0*5/State%CP%Accuracy%AccuracyBoost*0.65
            else
                ep = ep0
            end if
        else
            ep = ep0
        end if
        ep = ep*2
        t = min(State%ThermoData%tight_tau, State%ThermoData%OpacityToTime(k/ep))
    end function

    subroutine MassiveNuApproxTimeFor(State, k, times)
        ! equations.f90:2088-2103
        type(CAMBdata), intent(in) :: State
        real(dl), intent(in) :: k
        real(c_double), intent(out) :: times(CB_MAX_NU)
        real(dl) :: a_massive, boost
        integer :: nu_i
        times = 0
        do nu_i = 1, State%CP%Nu_mass_eigenstates
            boost = MassiveNuTransferLAccuracyBoost(State%nu_masses(nu_i), k, nu_tau_nonrelativistic_physical(nu_i))
            times(nu_i) = nu_tau_massive(nu_i)
            a_massive = 20000*k/State%nu_masses(nu_i)*MassiveNuTransferSwitchBoost(State%nu_masses(nu_i))*boost
            if (a_massive >= 0.99) then
                times(nu_i) = State%tau0 + 1
            else if (a_massive > 17.d0/State%nu_masses(nu_i)*MassiveNuTransferSwitchBoost(State%nu_masses(nu_i))) then
                times(nu_i) = max(times(nu_i), State%DeltaTime(0._dl, a_massive, 0.01_dl))
            end if
        end do
    end subroutine

    ! ---- CAMBdata -> cb_model (SURVEY section 8(b) "Inputs to marshal"); pointers are c_loc of the reference's allocatables ----

    subroutine FillModel(State, TimeSteps, n_src, gm)
        type(CAMBdata), intent(in), target :: State
        type(TRanges), intent(in), target :: TimeSteps          ! State%TimeSteps
        integer, intent(in) :: n_src                            ! ThisSources%SourceNum, cmbmain.f90:755-773
        type(cb_model), intent(out) :: gm
        type(CAMBparams), pointer :: P
        integer :: i, nw
        P => State%CP
        ! CAMBdata scalars, results.f90:184-227
        gm%grhog = State%grhog; gm%grhor = State%grhor; gm%grhornomass = State%grhornomass; gm%grhob = State%grhob
        gm%grhoc = State%grhoc; gm%grhov = State%grhov; gm%grhok = State%grhok
        gm%adotrad = State%adotrad; gm%tau0 = State%tau0; gm%tau_maxvis = State%tau_maxvis
        gm%taurst = State%ThermoData%tau_rst; gm%taurend = State%ThermoData%tau_rend                  ! cmbmain's taurst / taurend
        gm%akthom = State%akthom; gm%fHe = State%fHe
        gm%H0 = P%H0; gm%TCMB = P%TCMB; gm%Yhe = P%Yhe; gm%omch2 = P%omch2; gm%ombh2 = P%ombh2
        gm%Nu_mass_eigenstates = P%Nu_mass_eigenstates; gm%Num_Nu_massive = P%Num_Nu_massive
        gm%grhormass = 0; gm%nu_masses = 0
        gm%grhormass(1:P%Nu_mass_eigenstates) = State%grhormass(1:P%Nu_mass_eigenstates)
        gm%nu_masses(1:P%Nu_mass_eigenstates) = State%nu_masses(1:P%Nu_mass_eigenstates)
        ! CAMBparams flags and accuracy, model.f90:36-196
        gm%WantCls = merge(1, 0, P%WantCls); gm%WantTransfer = merge(1, 0, P%WantTransfer); gm%Want_CMB = merge(1, 0, P%Want_CMB)
        gm%WantTensors = merge(1, 0, P%WantTensors); gm%WantScalars = merge(1, 0, P%WantScalars); gm%DoLensing = merge(1, 0, P%DoLensing)
        gm%DoLateRadTruncation = merge(1, 0, P%DoLateRadTruncation); gm%Scalar_initial_condition = P%Scalar_initial_condition
        gm%MassiveNuMethod = P%MassiveNuMethod
        gm%transfer_high_precision = merge(1, 0, P%Transfer%high_precision)
        gm%transfer_accurate_massive_neutrinos = merge(1, 0, P%Transfer%accurate_massive_neutrinos)
        gm%AccuratePolarization = merge(1, 0, P%Accuracy%AccuratePolarization)
        gm%AccurateReionization = merge(1, 0, P%Accuracy%AccurateReionization)
        gm%AccuracyTarget = AccuracyTarget                                                             ! config.f90:12
        gm%AccuracyBoost = P%Accuracy%AccuracyBoost; gm%lAccuracyBoost = P%Accuracy%lAccuracyBoost
        gm%IntTolBoost = P%Accuracy%IntTolBoost; gm%TimeSwitchBoost = P%Accuracy%TimeSwitchBoost
        gm%BessIntBoost = P%Accuracy%BessIntBoost; gm%LimberBoost = P%Accuracy%LimberBoost; gm%Alens = P%Alens
        ! dark energy, DarkEnergyInterface.f90:85-105, 200-268: 0 = Lambda / fluid, 1 = PPF (cambb200_params.h CB_DE_*)
        gm%de_is_cosmological_constant = merge(1, 0, P%DarkEnergy%is_cosmological_constant)
        gm%de_num_perturb_equations = P%DarkEnergy%num_perturb_equations
        gm%de_model = 0; gm%de_w = -1; gm%de_wa = 0; gm%de_cs2 = 1; gm%de_c_gamma_ppf = 0.4_dl
        select type (DE => P%DarkEnergy)
        class is (TDarkEnergyEqnOfState)
            if (DE%use_tabulated_w) call GlobalError('GPU Boltzmann path: tabulated w(a) is not supported', error_unsupported_params)
            gm%de_w = DE%w_lam; gm%de_wa = DE%wa; gm%de_cs2 = DE%cs2_lam
        class default
            if (.not. P%DarkEnergy%is_cosmological_constant) &
                call GlobalError('GPU Boltzmann path: this dark energy model is not supported', error_unsupported_params)
        end select
        select type (DE => P%DarkEnergy)
        type is (TDarkEnergyPPF)
            gm%de_model = 1; gm%de_c_gamma_ppf = DE%c_Gamma_ppf
        end select
        ! TThermoData, results.f90:71-101
        gm%nthermo = State%ThermoData%nthermo; gm%tauminn = State%ThermoData%tauminn; gm%dlntau = State%ThermoData%dlntau
        gm%tight_tau = State%ThermoData%tight_tau; gm%matter_verydom_tau = State%ThermoData%matter_verydom_tau
        gm%thermo_horner = c_loc(State%ThermoData%thermo_horner); gm%cs2 = c_loc(State%ThermoData%cs2)
        if (thermo_tables_on_device) gm%thermo_horner = c_null_ptr   ! the library derives the tables from cs2, dotmu, emmu, adot, ScaleFactor
        gm%dotmu = c_loc(State%ThermoData%dotmu); gm%ddotmu = c_loc(State%ThermoData%ddotmu)
        gm%dddotmu = c_loc(State%ThermoData%dddotmu); gm%ddddotmu = c_loc(State%ThermoData%ddddotmu)
        gm%emmu = c_loc(State%ThermoData%emmu); gm%demmu = c_loc(State%ThermoData%demmu)
        gm%adot = c_loc(State%ThermoData%adot); gm%dadot = c_loc(State%ThermoData%dadot)               ! tensor pass only
        ! a / tau (results.f90:2441): read only when thermo_horner is c_null_ptr, i.e. when ThermoData%Init skips its packing loops
        ! (results.f90:2430-2494) and leaves the derivative tables and Horner coefficients to the device
        gm%scale_factor = c_loc(State%ThermoData%ScaleFactor)
        ! TimeSteps (TRanges, RangeUtils.f90:13-47)
        gm%n_tau = TimeSteps%npoints; gm%tau = c_loc(TimeSteps%points); gm%dtau = c_loc(TimeSteps%dpoints)
        if (allocated(tau_regions)) deallocate(tau_regions)
        allocate(tau_regions(TimeSteps%count))
        do i = 1, TimeSteps%count
            tau_regions(i) = cb_range(TimeSteps%R(i)%start_index, TimeSteps%R(i)%steps, merge(1, 0, TimeSteps%R(i)%IsLog), &
                TimeSteps%R(i)%Low, TimeSteps%R(i)%High, TimeSteps%R(i)%delta)
        end do
        gm%n_tau_regions = TimeSteps%count; gm%tau_regions = c_loc(tau_regions)
        gm%tau_lowest = TimeSteps%Lowest; gm%tau_highest = TimeSteps%Highest
        ! massive-neutrino momentum grid and switch times, massive_neutrinos.f90:120-178, equations.f90:197-198, 558-581
        gm%nqmax = State%NuPerturbations%nqmax
        gm%nu_q = 0; gm%nu_int_kernel = 0; gm%nu_tau_notmassless = 0
        gm%nu_tau_nonrelativistic = 0; gm%nu_tau_nonrelativistic_physical = 0; gm%nu_tau_massive = 0
        if (P%Num_Nu_massive > 0) then
            gm%nu_q(1:gm%nqmax) = State%NuPerturbations%nu_q(1:gm%nqmax)
            gm%nu_int_kernel(1:gm%nqmax) = State%NuPerturbations%nu_int_kernel(1:gm%nqmax)
            do i = 1, P%Nu_mass_eigenstates                                   ! (nqmax, max_nu) column-major, padded to CB_MAX_NQ rows
                gm%nu_tau_notmassless((i - 1)*CB_MAX_NQ + 1:(i - 1)*CB_MAX_NQ + gm%nqmax) = nu_tau_notmassless(1:gm%nqmax, i)
            end do
            gm%nu_tau_nonrelativistic(1:P%Nu_mass_eigenstates) = nu_tau_nonrelativistic(1:P%Nu_mass_eigenstates)
            gm%nu_tau_nonrelativistic_physical(1:P%Nu_mass_eigenstates) = nu_tau_nonrelativistic_physical(1:P%Nu_mass_eigenstates)
            gm%nu_tau_massive(1:P%Nu_mass_eigenstates) = nu_tau_massive(1:P%Nu_mass_eigenstates)
        end if
        ! transfer output times, results.f90 (State%Transfer_Times)
        gm%num_transfer_redshifts = 0; gm%transfer_times = c_null_ptr
        if (P%WantTransfer) then
            gm%num_transfer_redshifts = State%num_transfer_redshifts; gm%transfer_times = c_loc(State%Transfer_Times)
        end if
        ! cmbmain module scalars, cmbmain.f90:96-105, 785-815 (module variables of CAMBmain: passed in through the State here)
        gm%max_etak_scalar = max_etak_scalar; gm%max_etak_tensor = max_etak_tensor; gm%qmin = qmin; gm%qmax = qmax
        gm%WantLateTime = merge(1, 0, WantLateTime); gm%n_src = n_src
        gm%InitialConditionVector = 0
        if (allocated(P%InitialConditionVector)) gm%InitialConditionVector(1:5) = P%InitialConditionVector(1:5)
        ! source windows (config 4): TRedWin arrays of SetTimeStepWindows, results.f90:2781-2984
        nw = State%num_redshiftwindows
        gm%num_redshiftwindows = nw; gm%num_extra_redshiftwindows = State%num_extra_redshiftwindows
        if (allocated(windows)) deallocate(windows)
        allocate(windows(max(1, nw)))
        do i = 1, nw
            associate (W => State%Redshift_W(i))
                if (W%kind /= window_counts .and. W%kind /= window_lensing) &
                    call GlobalError('GPU Boltzmann path: 21cm source windows are not supported', error_unsupported_params)
                windows(i)%kind = W%kind; windows(i)%mag_index = W%mag_index
                windows(i)%Fq = W%Fq; windows(i)%dlog10Ndm = W%Window%dlog10Ndm; windows(i)%bias = W%Window%GetBias(1._dl, 1._dl)
                windows(i)%tau_start = W%tau_start; windows(i)%tau_end = W%tau_end; windows(i)%chi0 = W%chi0; windows(i)%sigma_tau = W%sigma_tau
                windows(i)%wing = c_loc(W%wing); windows(i)%wing2 = c_loc(W%wing2); windows(i)%dwing2 = c_loc(W%dwing2)
                windows(i)%ddwing2 = c_loc(W%ddwing2); windows(i)%wingtau = c_loc(W%wingtau); windows(i)%dwingtau = c_loc(W%dwingtau)
                windows(i)%winV = c_loc(W%winV); windows(i)%winF = c_loc(W%winF); windows(i)%win_lens = c_loc(W%win_lens)
                windows(i)%comoving_density_ev = c_loc(W%comoving_density_ev)
            end associate
        end do
        gm%windows = c_loc(windows)
        gm%SourceTerms%limber_windows = merge(1, 0, P%SourceTerms%limber_windows); gm%SourceTerms%limber_phi_lmin = P%SourceTerms%limber_phi_lmin
        gm%SourceTerms%counts_density = merge(1, 0, P%SourceTerms%counts_density); gm%SourceTerms%counts_redshift = merge(1, 0, P%SourceTerms%counts_redshift)
        gm%SourceTerms%counts_lensing = merge(1, 0, P%SourceTerms%counts_lensing); gm%SourceTerms%counts_velocity = merge(1, 0, P%SourceTerms%counts_velocity)
        gm%SourceTerms%counts_radial = merge(1, 0, P%SourceTerms%counts_radial); gm%SourceTerms%counts_timedelay = merge(1, 0, P%SourceTerms%counts_timedelay)
        gm%SourceTerms%counts_ISW = merge(1, 0, P%SourceTerms%counts_ISW); gm%SourceTerms%counts_potential = merge(1, 0, P%SourceTerms%counts_potential)
        gm%SourceTerms%counts_evolve = merge(1, 0, P%SourceTerms%counts_evolve)
        gm%tau_start_redshiftwindows = State%ThermoData%tau_start_redshiftwindows
        gm%Want_CMB_lensing = merge(1, 0, P%Want_CMB_lensing)
        gm%max_bessels_l_index = max_bessels_l_index                                                   ! set by GetLimberTransfers' index part, cmbmain.f90:420-470
        gm%Limber_l_min = 0
        ! things the device path does not implement are errors, never a fallback
        if (.not. State%flat) call GlobalError('GPU Boltzmann path: non-flat models are not supported', error_unsupported_params)
        if (P%WantVectors) call GlobalError('GPU Boltzmann path: vector modes are not supported', error_unsupported_params)
        if (P%Do21cm .or. P%Evolve_delta_xe .or. P%Evolve_delta_Ts) &
            call GlobalError('GPU Boltzmann path: 21cm / perturbed recombination are not supported', error_unsupported_params)
        if (P%CustomSources%num_custom_sources > 0) &
            call GlobalError('GPU Boltzmann path: custom sources are not supported', error_unsupported_params)
    end subroutine

    ! ---- loop A (cmbmain.f90:205-209): every source k-mode of ThisSources%Evolve_q ----
    subroutine GpuCalcSources(State, Evolve_q, n_src, LinearSrc)
        type(CAMBdata), intent(inout), target :: State
        type(TRanges), intent(in), target :: Evolve_q
        integer, intent(in) :: n_src
        real(dl), intent(inout), target :: LinearSrc(:, :, :)     ! (n_k, n_src, n_tau), cmbmain.f90:777
        type(cb_model) :: gm
        type(cb_modes) :: modes(1)
        type(c_ptr) :: dms(1), srcs(1), trs(1), sts(1)
        integer :: q_ix, n_k
        call GpuBoltzmann_Init()
        if (global_error_flag /= 0) return
        n_k = Evolve_q%npoints
        if (allocated(tight)) deallocate(tight, nuapp)
        allocate(tight(n_k), nuapp(CB_MAX_NU, n_k))
        do q_ix = 1, n_k
            tight(q_ix) = TightSwitchoffTimeFor(State, Evolve_q%points(q_ix))
            call MassiveNuApproxTimeFor(State, Evolve_q%points(q_ix), nuapp(:, q_ix))
        end do
        call FillModel(State, State%TimeSteps, n_src, gm)
        if (global_error_flag /= 0) return
        if (c_associated(dm)) call cb_model_destroy(ctx, dm)
        call Check(cb_model_create(ctx, gm, dm), 'cb_model_create')
        modes(1) = cb_modes(n_k, c_loc(Evolve_q%points), c_loc(tight), c_loc(nuapp), 0, 0, c_null_ptr)
        dms(1) = dm; srcs(1) = c_loc(LinearSrc); sts(1) = c_null_ptr; trs(1) = c_null_ptr
        if (State%CP%WantTransfer) trs(1) = c_loc(State%MT%TransferData)     ! rows 1:n_k of TransferData(13, num_q_trans, n_z)
        call Check(cb_evolve(ctx, 1, dms, modes, srcs, trs, sts), 'cb_evolve')
    end subroutine

    ! ---- loop A' (cmbmain.f90:1147-1163): transfer-only modes above the CMB grid ----
    subroutine GpuTransferOut(State, n_source_points)
        type(CAMBdata), intent(inout), target :: State
        integer, intent(in) :: n_source_points
        type(cb_modes) :: modes(1)
        type(c_ptr) :: dms(1), srcs(1), trs(1), sts(1)
        real(c_float), allocatable, target :: buf(:, :, :)
        integer :: n_extra, q_ix
        n_extra = State%MT%num_q_trans - n_source_points
        if (n_extra <= 0) return
        if (allocated(tight)) deallocate(tight, nuapp)
        allocate(tight(n_extra), nuapp(CB_MAX_NU, n_extra))
        do q_ix = 1, n_extra
            tight(q_ix) = TightSwitchoffTimeFor(State, State%MT%q_trans(n_source_points + q_ix))
            call MassiveNuApproxTimeFor(State, State%MT%q_trans(n_source_points + q_ix), nuapp(:, q_ix))
        end do
        allocate(buf(Transfer_max, n_extra, State%num_transfer_redshifts))
        modes(1) = cb_modes(n_extra, c_loc(State%MT%q_trans(n_source_points + 1)), c_loc(tight), c_loc(nuapp), 1, 0, c_null_ptr)
        dms(1) = dm; srcs(1) = c_null_ptr; trs(1) = c_loc(buf); sts(1) = c_null_ptr
        call Check(cb_evolve(ctx, 1, dms, modes, srcs, trs, sts), 'cb_evolve (TransferOut)')
        State%MT%TransferData(:, n_source_points + 1:State%MT%num_q_trans, :) = buf
    end subroutine

    ! ---- InitSourceInterpolation + loop B (cmbmain.f90:262-291) ----
    ! ---- MakeNonlinearSources (cmbmain.f90:1190-1247) on the device: Transfer_CacheNonLinRatios stays on the host (halofit is
    !      CP%NonLinearModel's business), only CAMB_PK%nonlin_ratio(n_k, n_z) goes down; the sources stay where they are.
    !      Called from cmbmain in place of `call MakeNonlinearSources` when use_gpu_boltzmann; GpuSourceToTransfers is then
    !      called with sources_rescaled = .false. ----
    subroutine GpuMakeNonlinearSources(State, Evolve_q)
        use NonLinear
        type(CAMBdata), intent(inout), target :: State
        type(TRanges), intent(in) :: Evolve_q
        if (allocated(State%CAMB_PK)) deallocate(State%CAMB_PK)
        call Transfer_CacheNonLinRatios(State)
        if (global_error_flag /= 0) return
        call Check(cb_make_nonlinear_sources(ctx, dm, Evolve_q%npoints, State%num_transfer_redshifts, &
            c_loc(State%CAMB_PK%nonlin_ratio), State%CP%NonLinearModel%Min_kh_nonlinear), 'cb_make_nonlinear_sources')
    end subroutine

    subroutine GpuSourceToTransfers(State, ThisCT, Evolve_q, ScaledSrc, sources_rescaled, BessRanges, bessel_horner)
        type(CAMBdata), intent(inout), target :: State
        type(ClTransferData), intent(inout), target :: ThisCT
        type(TRanges), intent(in), target :: Evolve_q, BessRanges
        real(dl), intent(in), target :: ScaledSrc(:, :, :), bessel_horner(:, :, :)   ! bessels.f90:20-41
        logical, intent(in) :: sources_rescaled                 ! MakeNonlinearSources ran on the host (cmbmain.f90:253-256)
        type(cb_bessel) :: gb
        type(c_ptr) :: dms(1), qs(1), dqs(1), deltas(1)
        integer(c_int) :: nq(1)
        integer :: i
        if (sources_rescaled) call Check(cb_set_sources(ctx, dm, Evolve_q%npoints, c_loc(Evolve_q%points), c_loc(ScaledSrc)), 'cb_set_sources')
        if (allocated(bess_regions)) deallocate(bess_regions)
        allocate(bess_regions(BessRanges%count))
        do i = 1, BessRanges%count
            bess_regions(i) = cb_range(BessRanges%R(i)%start_index, BessRanges%R(i)%steps, merge(1, 0, BessRanges%R(i)%IsLog), &
                BessRanges%R(i)%Low, BessRanges%R(i)%High, BessRanges%R(i)%delta)
        end do
        gb%n_l = ThisCT%max_index_nonlimber; gb%l = c_loc(ThisCT%ls%l); gb%nx = BessRanges%npoints; gb%x = c_loc(BessRanges%points)
        gb%n_regions = BessRanges%count; gb%regions = c_loc(bess_regions); gb%horner = c_loc(bessel_horner); gb%n_l_all = ThisCT%ls%nl
        if (build_bessel_on_device) then
            ! GenerateBessels skipped on the host (bessels.f90:140-228): the table is generated on the device from the l list and
            ! the x grid; bessel_horner is not read
            gb%horner = c_null_ptr
            call Check(cb_build_bessel(ctx, gb, c_null_ptr), 'cb_build_bessel')
        else
            call Check(cb_set_bessel(ctx, gb), 'cb_set_bessel')  ! the library re-uploads only when the table changed
        end if
        dms(1) = dm; nq(1) = ThisCT%q%npoints; qs(1) = c_loc(ThisCT%q%points); dqs(1) = c_loc(ThisCT%q%dpoints)
        deltas(1) = c_loc(ThisCT%Delta_p_l_k)                    ! (n_src, max_index_nonlimber, n_q), results.f90:3118-3119
        call Check(cb_project(ctx, 1, dms, nq, qs, dqs, deltas), 'cb_project')
    end subroutine

    ! ---- loop C (cmbmain.f90:322-323): CalcLimberScalCls + CalcScalCls ----
    subroutine GpuCalcScalCls(State, CTrans, iCl_scalar, iCl_Array)
        type(CAMBdata), intent(inout), target :: State
        type(ClTransferData), intent(in), target :: CTrans
        real(dl), intent(inout), target :: iCl_scalar(:, :)
        real(dl), intent(inout), target, optional :: iCl_Array(:, :, :)
        type(c_ptr) :: dms(1), pp(1), lp(1), icl(1), arr(1)
        integer :: q_ix, n, ell
        real(dl) :: chi
        if (allocated(pows)) deallocate(pows)
        allocate(pows(CTrans%q%npoints))
        do q_ix = 1, CTrans%q%npoints
            pows(q_ix) = State%CP%InitPower%ScalarPower(CTrans%q%points(q_ix))        ! cmbmain.f90:3199
        end do
        dms(1) = dm; pp(1) = c_loc(pows); icl(1) = c_loc(iCl_scalar); arr(1) = c_null_ptr
        if (present(iCl_Array)) arr(1) = c_loc(iCl_Array)
        if (State%num_redshiftwindows == 0) then
            call Check(cb_cl(ctx, 1, dms, 0, pp, icl), 'cb_cl')
            if (present(iCl_Array) .and. CTrans%NumSources > 2) call Check(cb_cl_array(ctx, 1, dms, pp, arr), 'cb_cl_array')
        else
            ! P_R on the Limber grid k = (l + 1/2)/chi_n (CalcLimberScalCls, cmbmain.f90:358-417)
            if (allocated(limb_pows)) deallocate(limb_pows)
            allocate(limb_pows(State%TimeSteps%npoints, CTrans%ls%nl))
            limb_pows = 0
            do ell = 1, CTrans%ls%nl
                do n = 1, State%TimeSteps%npoints
                    chi = State%tau0 - State%TimeSteps%points(n)
                    if (chi > 0) limb_pows(n, ell) = State%CP%InitPower%ScalarPower((CTrans%ls%l(ell) + 0.5_dl)/chi)
                end do
            end do
            lp(1) = c_loc(limb_pows)
            call Check(cb_cl_windows(ctx, 1, dms, pp, lp, icl, arr), 'cb_cl_windows')
        end if
    end subroutine

    ! ---- tensor pass of CAMB_GetResults (camb.f90:85-104): CalcTensCls, cmbmain.f90:3424-3476 ----
    subroutine GpuCalcTensCls(State, CTrans, iCl_tensor)
        type(CAMBdata), intent(inout), target :: State
        type(ClTransferData), intent(in), target :: CTrans
        real(dl), intent(inout), target :: iCl_tensor(:, :)
        type(c_ptr) :: dms(1), pp(1), icl(1)
        integer :: q_ix
        if (allocated(pows)) deallocate(pows)
        allocate(pows(CTrans%q%npoints))
        do q_ix = 1, CTrans%q%npoints
            pows(q_ix) = State%CP%InitPower%TensorPower(CTrans%q%points(q_ix))
        end do
        dms(1) = dm; pp(1) = c_loc(pows); icl(1) = c_loc(iCl_tensor)
        call Check(cb_cl(ctx, 1, dms, 1, pp, icl), 'cb_cl (tensor)')
    end subroutine

end module GpuBoltzmann
