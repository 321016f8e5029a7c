!> B200 engine as a third implementation of SPRAL's subtree plug-in interface
!> (symbolic_subtree_base / numeric_subtree_base, src/ssids/subtree.f90:26-126).
!>
!> Drop-in for src/ssids/gpu/subtree.f90: same module name and constructor
!> (construct_gpu_symbolic_subtree, called from src/ssids/anal.F90:1092-1095), so
!> replacing that file (and linking libspral_ssids_b200.so instead of the
!> src/ssids/gpu/kernels and src/cuda objects) switches SSIDS' GPU path to this
!> engine.  Everything numerical happens behind the C ABI of
!> include/spral_ssids_b200.h; this file only marshals arguments, exactly like
!> src/ssids/cpu/subtree.f90 does for the CPU engine.
!>
!> NOTE: no Fortran compiler exists in the build environment of this repository;
!> this file is kept deliberately thin so that it can be reviewed by reading.
module spral_ssids_gpu_subtree
   use, intrinsic :: iso_c_binding
   use spral_ssids_contrib, only : contrib_type
   use spral_ssids_datatypes
   use spral_ssids_inform, only : ssids_inform
   use spral_ssids_subtree, only : symbolic_subtree_base, numeric_subtree_base
   implicit none

   private
   public :: gpu_symbolic_subtree, construct_gpu_symbolic_subtree
   public :: gpu_numeric_subtree, gpu_free_contrib

   !> struct spral_ssids_b200_options == cpu_factor_options (src/ssids/cpu/cpu_iface.f90:21-31)
   type, bind(C) :: b200_options
      integer(C_INT) :: print_level
      logical(C_BOOL) :: action
      real(C_DOUBLE) :: small, u, multiplier
      integer(C_INT64_T) :: small_subtree_threshold
      integer(C_INT) :: cpu_block_size, pivot_method, failed_pivot_method
   end type b200_options

   !> struct spral_ssids_b200_stats == cpu_factor_stats (cpu_iface.f90:39-51) + cuda_error
   type, bind(C) :: b200_stats
      integer(C_INT) :: flag, num_delay
      integer(C_INT64_T) :: num_factor, num_flops
      integer(C_INT) :: num_neg, num_two, num_zero, maxfront, maxsupernode
      integer(C_INT) :: not_first_pass, not_second_pass, cuda_error
   end type b200_stats

   type, extends(symbolic_subtree_base) :: gpu_symbolic_subtree
      integer :: n
      type(C_PTR) :: csubtree = C_NULL_PTR
   contains
      procedure :: factor
      procedure :: cleanup => symbolic_cleanup
   end type gpu_symbolic_subtree

   type, extends(numeric_subtree_base) :: gpu_numeric_subtree
      logical(C_BOOL) :: posdef
      type(gpu_symbolic_subtree), pointer :: symbolic
      type(C_PTR) :: csubtree = C_NULL_PTR
   contains
      procedure :: get_contrib
      procedure :: solve_fwd
      procedure :: solve_diag
      procedure :: solve_diag_bwd
      procedure :: solve_bwd
      procedure :: enquire_posdef
      procedure :: enquire_indef
      procedure :: alter
      procedure :: cleanup => numeric_cleanup
   end type gpu_numeric_subtree

   interface
      type(C_PTR) function c_create_symbolic(device, n, sa, en, sptr, sparent, rptr, rlist, nptr, nlist, &
            ncontrib, contrib_idx, options) bind(C, name="spral_ssids_gpu_create_symbolic_subtree")
         import :: C_PTR, C_INT, C_INT64_T, b200_options
         integer(C_INT), value :: device, n, sa, en, ncontrib
         integer(C_INT), dimension(*), intent(in) :: sptr, sparent, rlist, contrib_idx
         integer(C_INT64_T), dimension(*), intent(in) :: rptr, nptr, nlist
         type(b200_options), intent(in) :: options
      end function c_create_symbolic
      subroutine c_destroy_symbolic(subtree) bind(C, name="spral_ssids_gpu_destroy_symbolic_subtree")
         import :: C_PTR
         type(C_PTR), value :: subtree
      end subroutine c_destroy_symbolic
      type(C_PTR) function c_create_numeric(posdef, symbolic, aval, scaling, child_contrib, options, stats) &
            bind(C, name="spral_ssids_gpu_create_num_subtree_dbl")
         import :: C_PTR, C_BOOL, C_DOUBLE, b200_options, b200_stats
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: symbolic, scaling
         real(C_DOUBLE), dimension(*), intent(in) :: aval
         type(C_PTR), dimension(*), intent(inout) :: child_contrib
         type(b200_options), intent(in) :: options
         type(b200_stats), intent(out) :: stats
      end function c_create_numeric
      subroutine c_destroy_numeric(posdef, subtree) bind(C, name="spral_ssids_gpu_destroy_num_subtree_dbl")
         import :: C_PTR, C_BOOL
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
      end subroutine c_destroy_numeric
      integer(C_INT) function c_solve_fwd(posdef, subtree, nrhs, x, ldx) bind(C, name="spral_ssids_gpu_subtree_solve_fwd_dbl")
         import :: C_PTR, C_BOOL, C_INT, C_DOUBLE
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
         integer(C_INT), value :: nrhs, ldx
         real(C_DOUBLE), dimension(*), intent(inout) :: x
      end function c_solve_fwd
      integer(C_INT) function c_solve_diag(posdef, subtree, nrhs, x, ldx) bind(C, name="spral_ssids_gpu_subtree_solve_diag_dbl")
         import :: C_PTR, C_BOOL, C_INT, C_DOUBLE
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
         integer(C_INT), value :: nrhs, ldx
         real(C_DOUBLE), dimension(*), intent(inout) :: x
      end function c_solve_diag
      integer(C_INT) function c_solve_diag_bwd(posdef, subtree, nrhs, x, ldx) &
            bind(C, name="spral_ssids_gpu_subtree_solve_diag_bwd_dbl")
         import :: C_PTR, C_BOOL, C_INT, C_DOUBLE
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
         integer(C_INT), value :: nrhs, ldx
         real(C_DOUBLE), dimension(*), intent(inout) :: x
      end function c_solve_diag_bwd
      integer(C_INT) function c_solve_bwd(posdef, subtree, nrhs, x, ldx) bind(C, name="spral_ssids_gpu_subtree_solve_bwd_dbl")
         import :: C_PTR, C_BOOL, C_INT, C_DOUBLE
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
         integer(C_INT), value :: nrhs, ldx
         real(C_DOUBLE), dimension(*), intent(inout) :: x
      end function c_solve_bwd
      subroutine c_enquire(posdef, subtree, piv_order, d) bind(C, name="spral_ssids_gpu_subtree_enquire_dbl")
         import :: C_PTR, C_BOOL
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree, piv_order, d
      end subroutine c_enquire
      subroutine c_alter(posdef, subtree, d) bind(C, name="spral_ssids_gpu_subtree_alter_dbl")
         import :: C_PTR, C_BOOL, C_DOUBLE
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
         real(C_DOUBLE), dimension(*), intent(in) :: d
      end subroutine c_alter
      subroutine c_get_contrib(posdef, subtree, n, val, ldval, rlist, ndelay, delay_perm, delay_val, lddelay) &
            bind(C, name="spral_ssids_gpu_subtree_get_contrib_dbl")
         import :: C_PTR, C_BOOL, C_INT
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
         integer(C_INT) :: n, ldval, ndelay, lddelay
         type(C_PTR) :: val, rlist, delay_perm, delay_val
      end subroutine c_get_contrib
      subroutine c_free_contrib(posdef, subtree) bind(C, name="spral_ssids_gpu_subtree_free_contrib_dbl")
         import :: C_PTR, C_BOOL
         logical(C_BOOL), value :: posdef
         type(C_PTR), value :: subtree
      end subroutine c_free_contrib
   end interface

contains

   subroutine copy_options_in(foptions, coptions)
      type(ssids_options), intent(in) :: foptions
      type(b200_options), intent(out) :: coptions
      coptions%print_level = foptions%print_level
      coptions%action = foptions%action
      coptions%small = foptions%small
      coptions%u = foptions%u
      coptions%multiplier = foptions%multiplier
      coptions%small_subtree_threshold = foptions%small_subtree_threshold
      coptions%cpu_block_size = foptions%cpu_block_size
      coptions%pivot_method = min(3, max(1, foptions%pivot_method))
      coptions%failed_pivot_method = min(2, max(1, foptions%failed_pivot_method))
   end subroutine copy_options_in

   !> Same accumulation as cpu_copy_stats_out (src/ssids/cpu/cpu_iface.f90:74-94).
   subroutine copy_stats_out(n, cstats, finform)
      integer, intent(in) :: n
      type(b200_stats), intent(in) :: cstats
      type(ssids_inform), intent(inout) :: finform
      if (cstats%flag .lt. 0) then
         finform%flag = min(finform%flag, cstats%flag)
         if (cstats%flag .eq. SSIDS_ERROR_CUDA_UNKNOWN) finform%cuda_error = cstats%cuda_error
      else
         finform%flag = max(finform%flag, cstats%flag)
      end if
      finform%maxfront = max(finform%maxfront, cstats%maxfront)
      finform%maxsupernode = max(finform%maxsupernode, cstats%maxsupernode)
      finform%num_delay = finform%num_delay + cstats%num_delay
      finform%num_factor = finform%num_factor + cstats%num_factor
      finform%num_flops = finform%num_flops + cstats%num_flops
      finform%num_neg = finform%num_neg + cstats%num_neg
      finform%num_two = finform%num_two + cstats%num_two
      finform%matrix_rank = finform%matrix_rank - cstats%num_zero
      finform%not_first_pass = finform%not_first_pass + cstats%not_first_pass
      finform%not_second_pass = finform%not_second_pass + cstats%not_second_pass
   end subroutine copy_stats_out

   !> Signature of the reference constructor (src/ssids/gpu/subtree.f90:76-90).
   function construct_gpu_symbolic_subtree(device, n, sa, en, sptr, sparent, rptr, rlist, nptr, nlist, &
         options) result(this)
      class(gpu_symbolic_subtree), pointer :: this
      integer, intent(in) :: device, n, sa, en
      integer, dimension(*), target, intent(in) :: sptr, sparent, rlist
      integer(long), dimension(*), target, intent(in) :: rptr, nptr
      integer(long), dimension(2,*), target, intent(in) :: nlist
      class(ssids_options), intent(in) :: options
      type(b200_options) :: coptions
      integer(C_INT) :: dummy(1)
      integer :: st
      nullify(this)
      allocate(this, stat=st)
      if (st .ne. 0) return
      this%n = n
      call copy_options_in(options, coptions)
      ! contributions from other parts are attached by the analyse phase through
      ! contrib_idx exactly as for cpu parts (anal.F90:1066-1097); none here = leaf part
      this%csubtree = c_create_symbolic(int(device, C_INT), int(n, C_INT), int(sa, C_INT), int(en, C_INT), &
         sptr, sparent, rptr, rlist, nptr, nlist, 0_C_INT, dummy, coptions)
   end function construct_gpu_symbolic_subtree

   subroutine symbolic_cleanup(this)
      class(gpu_symbolic_subtree), intent(inout) :: this
      call c_destroy_symbolic(this%csubtree)
      this%csubtree = C_NULL_PTR
   end subroutine symbolic_cleanup

   function factor(this, posdef, aval, child_contrib, options, inform, scaling)
      class(numeric_subtree_base), pointer :: factor
      class(gpu_symbolic_subtree), target, intent(inout) :: this
      logical, intent(in) :: posdef
      real(wp), dimension(*), target, intent(in) :: aval
      type(contrib_type), dimension(:), target, intent(inout) :: child_contrib
      type(ssids_options), intent(in) :: options
      type(ssids_inform), intent(inout) :: inform
      real(wp), dimension(*), target, optional, intent(in) :: scaling

      type(gpu_numeric_subtree), pointer :: sub
      type(b200_options) :: coptions
      type(b200_stats) :: cstats
      type(C_PTR) :: cscaling
      type(C_PTR), dimension(:), allocatable :: contrib_ptr
      integer :: i, st

      nullify(factor)
      allocate(sub, stat=st)
      if (st .ne. 0) then
         inform%flag = SSIDS_ERROR_ALLOCATION
         inform%stat = st
         return
      end if
      sub%symbolic => this
      sub%posdef = posdef
      allocate(contrib_ptr(max(1, size(child_contrib))))
      do i = 1, size(child_contrib)
         contrib_ptr(i) = C_LOC(child_contrib(i))
      end do
      cscaling = C_NULL_PTR
      if (present(scaling)) cscaling = C_LOC(scaling)
      call copy_options_in(options, coptions)
      sub%csubtree = c_create_numeric(sub%posdef, this%csubtree, aval, cscaling, contrib_ptr, coptions, cstats)
      call copy_stats_out(this%n, cstats, inform)
      factor => sub
   end function factor

   subroutine numeric_cleanup(this)
      class(gpu_numeric_subtree), intent(inout) :: this
      call c_destroy_numeric(this%posdef, this%csubtree)
      this%csubtree = C_NULL_PTR
   end subroutine numeric_cleanup

   function get_contrib(this)
      type(contrib_type) :: get_contrib
      class(gpu_numeric_subtree), intent(in) :: this
      type(C_PTR) :: cval, crlist, delay_perm, delay_val
      call c_get_contrib(this%posdef, this%csubtree, get_contrib%n, cval, get_contrib%ldval, crlist, &
         get_contrib%ndelay, delay_perm, delay_val, get_contrib%lddelay)
      call c_f_pointer(cval, get_contrib%val, shape=(/ get_contrib%n**2 /))
      call c_f_pointer(crlist, get_contrib%rlist, shape=(/ get_contrib%n /))
      if (c_associated(delay_val)) then
         call c_f_pointer(delay_perm, get_contrib%delay_perm, shape=(/ get_contrib%ndelay /))
         call c_f_pointer(delay_val, get_contrib%delay_val, shape=(/ get_contrib%ndelay*get_contrib%lddelay /))
      else
         nullify(get_contrib%delay_perm)
         nullify(get_contrib%delay_val)
      end if
      get_contrib%owner = 1      ! contrib_free (src/ssids/contrib_free.f90:17-31) dispatches to gpu_free_contrib
      get_contrib%posdef = this%posdef
      get_contrib%owner_ptr = this%csubtree
   end function get_contrib

   subroutine gpu_free_contrib(posdef, csubtree)
      logical(C_BOOL), intent(in) :: posdef
      type(C_PTR), intent(inout) :: csubtree
      call c_free_contrib(posdef, csubtree)
   end subroutine gpu_free_contrib

   subroutine solve_fwd(this, nrhs, x, ldx, inform)
      class(gpu_numeric_subtree), intent(inout) :: this
      integer, intent(in) :: nrhs
      real(wp), dimension(*), intent(inout) :: x
      integer, intent(in) :: ldx
      type(ssids_inform), intent(inout) :: inform
      integer(C_INT) :: flag
      flag = c_solve_fwd(this%posdef, this%csubtree, nrhs, x, ldx)
      if (flag .ne. SSIDS_SUCCESS) inform%flag = flag
   end subroutine solve_fwd

   subroutine solve_diag(this, nrhs, x, ldx, inform)
      class(gpu_numeric_subtree), intent(inout) :: this
      integer, intent(in) :: nrhs
      real(wp), dimension(*), intent(inout) :: x
      integer, intent(in) :: ldx
      type(ssids_inform), intent(inout) :: inform
      integer(C_INT) :: flag
      flag = c_solve_diag(this%posdef, this%csubtree, nrhs, x, ldx)
      if (flag .ne. SSIDS_SUCCESS) inform%flag = flag
   end subroutine solve_diag

   subroutine solve_diag_bwd(this, nrhs, x, ldx, inform)
      class(gpu_numeric_subtree), intent(inout) :: this
      integer, intent(in) :: nrhs
      real(wp), dimension(*), intent(inout) :: x
      integer, intent(in) :: ldx
      type(ssids_inform), intent(inout) :: inform
      integer(C_INT) :: flag
      flag = c_solve_diag_bwd(this%posdef, this%csubtree, nrhs, x, ldx)
      if (flag .ne. SSIDS_SUCCESS) inform%flag = flag
   end subroutine solve_diag_bwd

   subroutine solve_bwd(this, nrhs, x, ldx, inform)
      class(gpu_numeric_subtree), intent(inout) :: this
      integer, intent(in) :: nrhs
      real(wp), dimension(*), intent(inout) :: x
      integer, intent(in) :: ldx
      type(ssids_inform), intent(inout) :: inform
      integer(C_INT) :: flag
      flag = c_solve_bwd(this%posdef, this%csubtree, nrhs, x, ldx)
      if (flag .ne. SSIDS_SUCCESS) inform%flag = flag
   end subroutine solve_bwd

   subroutine enquire_posdef(this, d)
      class(gpu_numeric_subtree), intent(in) :: this
      real(wp), dimension(*), target, intent(out) :: d
      call c_enquire(this%posdef, this%csubtree, C_NULL_PTR, C_LOC(d))
   end subroutine enquire_posdef

   subroutine enquire_indef(this, piv_order, d)
      class(gpu_numeric_subtree), intent(in) :: this
      integer, dimension(*), target, optional, intent(out) :: piv_order
      real(wp), dimension(2,*), target, optional, intent(out) :: d
      type(C_PTR) :: dptr, poptr
      poptr = C_NULL_PTR
      if (present(piv_order)) poptr = C_LOC(piv_order)
      dptr = C_NULL_PTR
      if (present(d)) dptr = C_LOC(d)
      call c_enquire(this%posdef, this%csubtree, poptr, dptr)
   end subroutine enquire_indef

   subroutine alter(this, d)
      class(gpu_numeric_subtree), target, intent(inout) :: this
      real(wp), dimension(2,*), intent(in) :: d
      call c_alter(this%posdef, this%csubtree, d)
   end subroutine alter

end module spral_ssids_gpu_subtree
