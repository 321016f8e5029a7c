!> Replacement for driver/cuda_helper_gpu.f90 (cuda_init, :9-43): device count and
!> warm-up through the thin C ABI instead of the Fortran bindings of src/cuda.
module cuda_helper
   use, intrinsic :: iso_c_binding
   implicit none
   private
   public :: cuda_init
   interface
      integer(C_INT) function c_cuda_init(cnt) bind(C, name="spral_ssids_b200_cuda_init")
         import :: C_INT
         integer(C_INT), intent(out) :: cnt
      end function c_cuda_init
   end interface
contains
   subroutine cuda_init(cnt)
      integer, intent(out) :: cnt
      integer(C_INT) :: ccnt, cuda_error
      cuda_error = c_cuda_init(ccnt)
      cnt = ccnt
      if (cuda_error .ne. 0) then
         print *, "CUDA initialisation failed with error ", cuda_error
         cnt = 0
      else
         print "(a,i3,a)", " Detected ", cnt, " GPU(s)"
      end if
   end subroutine cuda_init
end module cuda_helper
