!>
!! @file m_b200_bindings.f90
!! @brief ISO_C_BINDING interface to libmfc_b200.so (include/mfc_b200.h) and the three
!!        wrapper routines the unchanged MicroFC host calls in place of its OpenACC hot path.
!!
!! Source-only artefact: this image has no Fortran compiler, so the file is shipped
!! uncompiled (INTEGRATION.md shows where it plugs into src/simulation).  Every type below
!! mirrors one declaration of include/mfc_b200.h member for member; tests/test_abi.py
!! checks the C side (sizes, offsets, exported symbols), and tests/test_fortran_binding.py
!! checks that this file names exactly the exported symbols with matching argument counts.
!!
!! Reference call sites replaced (paths relative to the MicroFC tree):
!!   p_main.fpp:131-151,175  module initialisers        -> s_b200_initialize
!!   p_main.fpp:188-193      !$acc update device(q)     -> s_b200_upload
!!   p_main.fpp:229-235      s_{1,2,3}_order_tvd_rk     -> s_b200_time_step
!!   p_main.fpp:218,296      !$acc update host(q)       -> s_b200_download
!!   p_main.fpp:329-341      module finalisers          -> s_b200_finalize
module m_b200_bindings

    use, intrinsic :: iso_c_binding

    implicit none

    private
    public :: mfc_b200_params_t, &
              s_b200_initialize, s_b200_upload, s_b200_time_step, &
              s_b200_download, s_b200_download_prim, s_b200_compute_rhs, s_b200_finalize, &
              mfc_b200_patch_t, s_b200_generate_initial_condition

    integer(c_int), parameter :: MFC_B200_MAX_FLUIDS = 4
    integer(c_int), parameter :: MFC_B200_ABI_VERSION = 1
    integer(c_int), parameter :: MFC_B200_MAX_PATCHES = 10

    !> mfc_b200_patch_t (include/mfc_b200.h): patch_icpp(i) as pre_process holds it
    !! (ic_patch_parameters, src/common/m_derived_types.f90:55-103)
    type, bind(C) :: mfc_b200_patch_t
        integer(c_int32_t) :: geometry
        integer(c_int32_t) :: smoothen
        integer(c_int32_t) :: smooth_patch_id
        integer(c_int32_t) :: alter_patch(0:MFC_B200_MAX_PATCHES)
        real(c_double)     :: x_centroid, y_centroid, z_centroid
        real(c_double)     :: length_x, length_y, length_z
        real(c_double)     :: radius
        real(c_double)     :: radii(3)
        real(c_double)     :: normal(3)
        real(c_double)     :: epsilon
        real(c_double)     :: smooth_coeff
        real(c_double)     :: vel(3)
        real(c_double)     :: pres
        real(c_double)     :: alpha_rho(MFC_B200_MAX_FLUIDS)
        real(c_double)     :: alpha(MFC_B200_MAX_FLUIDS)
    end type mfc_b200_patch_t

    !> mfc_b200_params_t (include/mfc_b200.h): everything the hot path reads from module
    !! globals in the reference (m_global_parameters.fpp:32-215), after decomposition.
    type, bind(C) :: mfc_b200_params_t
        integer(c_int32_t) :: abi_version
        integer(c_int32_t) :: m, n, p
        integer(c_int32_t) :: m_glb, n_glb, p_glb
        integer(c_int32_t) :: num_dims
        integer(c_int32_t) :: num_fluids
        integer(c_int32_t) :: sys_size
        integer(c_int32_t) :: buff_size
        integer(c_int32_t) :: weno_order
        real(c_double)     :: weno_eps
        integer(c_int32_t) :: time_stepper
        integer(c_int32_t) :: weno_Re_flux
        integer(c_int32_t) :: run_time_info
        integer(c_int32_t) :: t_step_start
        integer(c_int32_t) :: t_step_stop
        integer(c_int32_t) :: bc(6)            !< x_beg, x_end, y_beg, y_end, z_beg, z_end
        integer(c_int32_t) :: proc_rank, num_procs
        integer(c_int32_t) :: proc_coords(3)
        integer(c_int32_t) :: num_procs_dir(3)
        real(c_double)     :: gammas(MFC_B200_MAX_FLUIDS)
        real(c_double)     :: pi_infs(MFC_B200_MAX_FLUIDS)
        real(c_double)     :: Re(2, MFC_B200_MAX_FLUIDS)   !< C: Re[fluid][2]
        type(c_ptr)        :: cb(3)
        type(c_ptr)        :: cc(3)
        type(c_ptr)        :: ds(3)
        integer(c_int32_t) :: strict_math
        integer(c_int32_t) :: device
        integer(c_int32_t) :: reserved(6)
    end type mfc_b200_params_t

    interface

        function mfc_b200_init(params) bind(C, name='mfc_b200_init') result(ierr)
            import :: c_int, mfc_b200_params_t
            type(mfc_b200_params_t), intent(in) :: params
            integer(c_int) :: ierr
        end function mfc_b200_init

        function mfc_b200_get_unique_id(id) bind(C, name='mfc_b200_get_unique_id') result(ierr)
            import :: c_int, c_char
            character(kind=c_char), intent(out) :: id(128)
            integer(c_int) :: ierr
        end function mfc_b200_get_unique_id

        function mfc_b200_comm_init(id, rank, nranks) bind(C, name='mfc_b200_comm_init') result(ierr)
            import :: c_int, c_char
            character(kind=c_char), intent(in) :: id(128)
            integer(c_int), value :: rank, nranks
            integer(c_int) :: ierr
        end function mfc_b200_comm_init

        function mfc_b200_upload(q_cons) bind(C, name='mfc_b200_upload') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: q_cons(*)      !< sys_size base pointers
            integer(c_int) :: ierr
        end function mfc_b200_upload

        function mfc_b200_step(t_step, dt, stab, step_seconds) bind(C, name='mfc_b200_step') result(ierr)
            import :: c_int, c_double
            integer(c_int), value :: t_step
            real(c_double), value :: dt
            real(c_double), intent(inout) :: stab(3)  !< ICFL max, VCFL max, Rc min
            real(c_double), intent(out) :: step_seconds
            integer(c_int) :: ierr
        end function mfc_b200_step

        function mfc_b200_step_async(t_step, dt, n_steps) bind(C, name='mfc_b200_step_async') result(ierr)
            import :: c_int, c_double
            integer(c_int), value :: t_step, n_steps
            real(c_double), value :: dt
            integer(c_int) :: ierr
        end function mfc_b200_step_async

        function mfc_b200_sync() bind(C, name='mfc_b200_sync') result(ierr)
            import :: c_int
            integer(c_int) :: ierr
        end function mfc_b200_sync

        function mfc_b200_compute_rhs(q_cons, rhs) bind(C, name='mfc_b200_compute_rhs') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: q_cons(*), rhs(*)
            integer(c_int) :: ierr
        end function mfc_b200_compute_rhs

        function mfc_b200_generate_initial_condition(num_patches, patches, cc, ds_min) &
            bind(C, name='mfc_b200_generate_initial_condition') result(ierr)
            import :: c_int, c_int32_t, c_ptr, c_double, mfc_b200_patch_t
            integer(c_int32_t), value :: num_patches
            type(mfc_b200_patch_t), intent(in) :: patches(*)
            type(c_ptr), intent(in) :: cc(3)
            real(c_double), value :: ds_min
            integer(c_int) :: ierr
        end function mfc_b200_generate_initial_condition

        function mfc_b200_generate_initial_condition2(num_patches, patches, cc, cb, ds_min) &
            bind(C, name='mfc_b200_generate_initial_condition2') result(ierr)
            import :: c_int, c_int32_t, c_ptr, c_double, mfc_b200_patch_t
            integer(c_int32_t), value :: num_patches
            type(mfc_b200_patch_t), intent(in) :: patches(*)
            type(c_ptr), intent(in) :: cc(3), cb(3)
            real(c_double), value :: ds_min
            integer(c_int) :: ierr
        end function mfc_b200_generate_initial_condition2

        function mfc_b200_download(q_cons) bind(C, name='mfc_b200_download') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: q_cons(*)
            integer(c_int) :: ierr
        end function mfc_b200_download

        function mfc_b200_download_prim(q_prim) bind(C, name='mfc_b200_download_prim') result(ierr)
            import :: c_int, c_ptr
            type(c_ptr), intent(in) :: q_prim(*)
            integer(c_int) :: ierr
        end function mfc_b200_download_prim

        function mfc_b200_finalize() bind(C, name='mfc_b200_finalize') result(ierr)
            import :: c_int
            integer(c_int) :: ierr
        end function mfc_b200_finalize

        function mfc_b200_last_error() bind(C, name='mfc_b200_last_error') result(msg)
            import :: c_ptr
            type(c_ptr) :: msg
        end function mfc_b200_last_error

        function c_strlen(s) bind(C, name='strlen') result(n)
            import :: c_ptr, c_size_t
            type(c_ptr), value :: s
            integer(c_size_t) :: n
        end function c_strlen

    end interface

contains

    !> Fail-stop like the reference (m_mpi_common.fpp:307-320): print the library's message
    !! and abort all ranks.
    subroutine s_b200_check(ierr, where)
        use m_mpi_common, only: s_mpi_abort
        integer(c_int), intent(in) :: ierr
        character(len=*), intent(in) :: where
        type(c_ptr) :: cmsg
        character(kind=c_char), pointer :: fmsg(:)
        integer :: i, n

        if (ierr == 0) return
        cmsg = mfc_b200_last_error()
        if (c_associated(cmsg)) then
            n = int(c_strlen(cmsg))
            call c_f_pointer(cmsg, fmsg, [n])
            print '(A)', 'libmfc_b200: '//where//': '//transfer(fmsg(1:n), repeat(' ', n))
        else
            print '(A,I0)', 'libmfc_b200: '//where//' failed with code ', ierr
        end if
        i = 0
        call s_mpi_abort()
    end subroutine s_b200_check

    !> Replaces the module initialisers of p_main.fpp:131-151,175.  Call after
    !! s_populate_grid_variables_buffers (p_main.fpp:170): the WENO coefficients need the
    !! ghosted x_cb / y_cb (m_weno.fpp:103-159).
    subroutine s_b200_initialize()
        use m_global_parameters
        use m_mpi_common, only: s_mpi_abort
#ifdef MFC_MPI
        use mpi
#endif
        type(mfc_b200_params_t) :: prm
        character(kind=c_char) :: nccl_id(128)
        integer :: i, ierr

        prm%abi_version = MFC_B200_ABI_VERSION
        prm%m = m; prm%n = n; prm%p = 0
        prm%m_glb = m_glb; prm%n_glb = n_glb; prm%p_glb = 0
        prm%num_dims = num_dims
        prm%num_fluids = num_fluids
        prm%sys_size = sys_size
        prm%buff_size = buff_size
        prm%weno_order = weno_order
        prm%weno_eps = weno_eps
        prm%time_stepper = time_stepper
        prm%weno_Re_flux = merge(1, 0, weno_Re_flux)
        prm%run_time_info = merge(1, 0, run_time_info)
        prm%t_step_start = t_step_start
        prm%t_step_stop = t_step_stop
        ! after s_mpi_decompose_computational_domain: < 0 physical code, >= 0 neighbour rank
        ! (m_mpi_proxy.fpp:242-255, 300-311)
        prm%bc = [bc_x%beg, bc_x%end, bc_y%beg, bc_y%end, -3, -3]
        prm%proc_rank = proc_rank
        prm%num_procs = num_procs
        prm%proc_coords = 0
        prm%proc_coords(1:num_dims) = proc_coords(1:num_dims)   ! m_global_parameters.fpp:122,406
        ! num_procs_x / num_procs_y are locals of s_mpi_decompose_computational_domain
        ! (m_mpi_proxy.fpp:138); informational only, the neighbour ranks in bc(:) are what the
        ! halo exchange uses
        prm%num_procs_dir = 0
        prm%gammas = 0d0; prm%pi_infs = 0d0; prm%Re = dflt_real
        do i = 1, num_fluids
            prm%gammas(i) = fluid_pp(i)%gamma      ! m_variables_conversion.fpp:253-260
            prm%pi_infs(i) = fluid_pp(i)%pi_inf
            prm%Re(:, i) = fluid_pp(i)%Re(:)
        end do
        ! ghosted metric arrays as allocated in m_global_parameters.fpp:386-394
        prm%cb = c_null_ptr; prm%cc = c_null_ptr; prm%ds = c_null_ptr
        prm%cb(1) = c_loc(x_cb(-1 - buff_size))
        prm%cc(1) = c_loc(x_cc(-buff_size))
        prm%ds(1) = c_loc(dx(-buff_size))
        if (n > 0) then
            prm%cb(2) = c_loc(y_cb(-1 - buff_size))
            prm%cc(2) = c_loc(y_cc(-buff_size))
            prm%ds(2) = c_loc(dy(-buff_size))
        end if
        prm%strict_math = 0
        prm%device = -1                            ! local_rank mod devNum, p_main.fpp:95-99
        prm%reserved = 0

        call s_b200_check(mfc_b200_init(prm), 'mfc_b200_init')

        if (num_procs > 1) then
            ! replaces MPI_CART_CREATE for the data path: NCCL communicator over NVLink
            if (proc_rank == 0) call s_b200_check(mfc_b200_get_unique_id(nccl_id), 'mfc_b200_get_unique_id')
#ifdef MFC_MPI
            call MPI_BCAST(nccl_id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
#endif
            call s_b200_check(mfc_b200_comm_init(nccl_id, proc_rank, num_procs), 'mfc_b200_comm_init')
        end if
    end subroutine s_b200_initialize

    !> p_main.fpp:188-193 -- "!$acc update device(q_cons_ts(1)%vf(i)%sf)"
    subroutine s_b200_upload(q_cons_vf)
        use m_derived_types
        use m_global_parameters, only: sys_size
        type(scalar_field), dimension(sys_size), intent(in), target :: q_cons_vf
        type(c_ptr) :: ptrs(sys_size)
        integer :: i
        do i = 1, sys_size
            ptrs(i) = c_loc(q_cons_vf(i)%sf)
        end do
        call s_b200_check(mfc_b200_upload(ptrs), 'mfc_b200_upload')
    end subroutine s_b200_upload

    !> Same signature as s_3rd_order_tvd_rk(t_step, time_avg) (m_time_steppers.fpp:271);
    !! time_stepper = 1, 2, 3 is selected inside the library.  dt is read from the module
    !! global at every call because the host mutates it (p_main.fpp:287).  The ICFL row of
    !! run_time.inf (m_data_output.fpp:296-305) is written by the unchanged host routine
    !! from icfl_max_glb.
    subroutine s_b200_time_step(t_step, time_avg)
        use m_global_parameters, only: dt, run_time_info, t_step_start, proc_rank
        use m_data_output, only: icfl_max_glb, vcfl_max_glb, Rc_min_glb
        integer, intent(in) :: t_step
        real(kind(0d0)), intent(inout) :: time_avg
        real(c_double) :: stab(3), secs

        stab = [icfl_max_glb, vcfl_max_glb, Rc_min_glb]
        call s_b200_check(mfc_b200_step(int(t_step, c_int), dt, stab, secs), 'mfc_b200_step')
        if (run_time_info) then
            icfl_max_glb = stab(1); vcfl_max_glb = stab(2); Rc_min_glb = stab(3)
        end if
        ! m_time_steppers.fpp:352-358: running mean over the steps after the 4th
        if (t_step >= 4) then
            time_avg = (abs(secs) + (t_step - 4)*time_avg)/(t_step - 3)
        else
            time_avg = 0d0
        end if
    end subroutine s_b200_time_step

    !> The pre_process stage on the device (SURVEY 8f-2): lays patch_icpp(1:num_patches) over this
    !! rank's cells exactly like s_generate_initial_condition (src/pre_process/
    !! m_initial_condition.fpp:42-113) and leaves the conservative state in HBM, replacing the
    !! restart-file round trip + s_b200_upload.  x_cc_pre/y_cc_pre: pre_process' cell centres
    !! (x_cb(i-1) + x_cb(i))/2 of the local cells (m_start_up.fpp:717,743); ds_min: the global
    !! minimum cell width (s_mpi_reduce_min, :720).
    !! x_cb_pre/y_cb_pre: the right cell boundaries x_cb(0:m), y_cb(0:n), which the analytical
    !! patches (geometry 7, 15; m_create_patches.fpp:466-467,:527-528) evaluate their bump at.
    subroutine s_b200_generate_initial_condition(patch_icpp, num_patches, x_cc_pre, y_cc_pre, x_cb_pre, y_cb_pre, ds_min)
        use m_derived_types
        type(ic_patch_parameters), intent(in) :: patch_icpp(:)
        integer, intent(in) :: num_patches
        real(kind(0d0)), intent(in), target :: x_cc_pre(0:), y_cc_pre(0:), x_cb_pre(0:), y_cb_pre(0:)
        real(kind(0d0)), intent(in) :: ds_min
        type(mfc_b200_patch_t) :: c(num_patches)
        type(c_ptr) :: cc(3), cb(3)
        integer :: i, k
        do i = 1, num_patches
            c(i)%geometry = patch_icpp(i)%geometry
            c(i)%smoothen = merge(1, 0, patch_icpp(i)%smoothen)
            c(i)%smooth_patch_id = patch_icpp(i)%smooth_patch_id
            c(i)%alter_patch = 0
            do k = 0, num_patches
                c(i)%alter_patch(k) = merge(1, 0, patch_icpp(i)%alter_patch(k))
            end do
            c(i)%x_centroid = patch_icpp(i)%x_centroid; c(i)%y_centroid = patch_icpp(i)%y_centroid
            c(i)%z_centroid = patch_icpp(i)%z_centroid
            c(i)%length_x = patch_icpp(i)%length_x; c(i)%length_y = patch_icpp(i)%length_y; c(i)%length_z = 0d0
            c(i)%radius = patch_icpp(i)%radius
            c(i)%radii = 0d0; c(i)%radii(1:2) = patch_icpp(i)%radii
            c(i)%normal = 0d0; c(i)%normal(1:2) = patch_icpp(i)%normal
            c(i)%epsilon = patch_icpp(i)%epsilon
            c(i)%smooth_coeff = patch_icpp(i)%smooth_coeff
            c(i)%vel = 0d0; c(i)%vel(1:2) = patch_icpp(i)%vel
            c(i)%pres = patch_icpp(i)%pres
            c(i)%alpha_rho = patch_icpp(i)%alpha_rho(1:MFC_B200_MAX_FLUIDS)
            c(i)%alpha = patch_icpp(i)%alpha(1:MFC_B200_MAX_FLUIDS)
        end do
        cc(1) = c_loc(x_cc_pre); cc(2) = c_loc(y_cc_pre); cc(3) = c_null_ptr
        cb(1) = c_loc(x_cb_pre); cb(2) = c_loc(y_cb_pre); cb(3) = c_null_ptr
        call s_b200_check(mfc_b200_generate_initial_condition2(int(num_patches, c_int32_t), c, cc, cb, ds_min), &
                          'mfc_b200_generate_initial_condition2')
    end subroutine s_b200_generate_initial_condition

    !> p_main.fpp:218,296 and m_time_steppers.fpp:374 -- "!$acc update host(...)"
    subroutine s_b200_download(q_cons_vf)
        use m_derived_types
        use m_global_parameters, only: sys_size
        type(scalar_field), dimension(sys_size), intent(inout), target :: q_cons_vf
        type(c_ptr) :: ptrs(sys_size)
        integer :: i
        do i = 1, sys_size
            ptrs(i) = c_loc(q_cons_vf(i)%sf)
        end do
        call s_b200_check(mfc_b200_download(ptrs), 'mfc_b200_download')
    end subroutine s_b200_download

    subroutine s_b200_download_prim(q_prim_vf)
        use m_derived_types
        use m_global_parameters, only: sys_size
        type(scalar_field), dimension(sys_size), intent(inout), target :: q_prim_vf
        type(c_ptr) :: ptrs(sys_size)
        integer :: i
        do i = 1, sys_size
            ptrs(i) = c_loc(q_prim_vf(i)%sf)
        end do
        call s_b200_check(mfc_b200_download_prim(ptrs), 'mfc_b200_download_prim')
    end subroutine s_b200_download_prim

    !> Same argument meaning as s_compute_rhs(q_cons_vf, q_prim_vf, rhs_vf, t_step)
    !! (m_rhs.fpp:405) for hosts that keep their own stepper: q_cons_vf ghosted fields in,
    !! rhs_vf fields of shape (0:m, 0:n) out.
    subroutine s_b200_compute_rhs(q_cons_vf, rhs_vf)
        use m_derived_types
        use m_global_parameters, only: sys_size
        type(scalar_field), dimension(sys_size), intent(in), target :: q_cons_vf
        type(scalar_field), dimension(sys_size), intent(inout), target :: rhs_vf
        type(c_ptr) :: qp(sys_size), rp(sys_size)
        integer :: i
        do i = 1, sys_size
            qp(i) = c_loc(q_cons_vf(i)%sf)
            rp(i) = c_loc(rhs_vf(i)%sf)
        end do
        call s_b200_check(mfc_b200_compute_rhs(qp, rp), 'mfc_b200_compute_rhs')
    end subroutine s_b200_compute_rhs

    !> p_main.fpp:329-341
    subroutine s_b200_finalize()
        call s_b200_check(mfc_b200_finalize(), 'mfc_b200_finalize')
    end subroutine s_b200_finalize

end module m_b200_bindings
