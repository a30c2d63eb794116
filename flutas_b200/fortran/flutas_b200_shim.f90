!
! flutas_b200_shim.f90 -- thin iso_c_binding shim between FluTAS's Fortran call sites and libflutas_b200.so.
!
! Build recipe (INTEGRATION.md): compile FluTAS WITHOUT _OPENACC (e.g. ARCH=generic-gnu), drop this file into src/,
! REMOVE fft.o solver_cpu.o solver_gpu.o fillps.o correc.o chkdiv.o from the object list (this file provides modules
! with the same names) and link with -lflutas_b200.  Module procedures carry the reference's names and argument
! lists, so the RK loop of src/apps/<APP>/main__<APP>.f90 (:693-749), initsolver.f90 and the BC-driven transform
! selection stay unchanged:
!
!   same module names (object files replaced, no source edit)
!     mod_fft        fftini, fftend          src/fft.f90:24,159          (initsolver.f90:117 calls fftini)
!     mod_solver_cpu solver_cpu              src/solver_cpu.f90:20       (main__single_phase.f90:716, non-_OPENACC build)
!     mod_solver_gpu solver_gpu              src/solver_gpu.f90:31       (main:714; needs lambdaxy in the CPU order)
!     mod_fillps     fillps                  src/fillps.f90:16
!     mod_correc     correc                  src/correc.f90:16
!     mod_chkdiv     chkdiv                  src/chkdiv.f90:18
!   next rows (SURVEY.md 8f): the reference modules hold other procedures too, so these carry a _b200 suffix and the
!   main's `use` line changes by one word -- or, for the single_phase app, compile with -DB200_REPLACE_MOD_BOUND
!   -DB200_REPLACE_MOD_CHKDT and drop bound.o / chkdt.o (main and sanity.f90 use nothing else from them)
!     mod_bound[_b200]  boundp, bounduvw, updt_rhs_b   src/bound.f90:146,17,829
!     mod_chkdt[_b200]  chkdt_sp                       src/chkdt.f90:117
!     mod_source_b200   pres_sp_src, pres_tw_src, pold_update   src/source.f90:311,247 ; main:693-699,734-740
!     mod_load_b200     load                           src/load.f90:21
!
! Fields: any array may stay an ordinary Fortran (host) array -- every entry point then stages it through the device and
! returns when the result is back (the `e2e` figure of bench.py).  For device-resident runs allocate the fields with
! b200_alloc_field (managed memory, like the reference's GPU build: main__single_phase.f90:157-163,268-299).
!
! Several ranks (dims_in = (1,N), one rank per GPU of one NVSwitch box): the first solver call exchanges the CUDA IPC
! handles of the peers' exchange buffers with one MPI_ALLGATHER, all-gathers the eigenvalue windows of initsolver once,
! and registers an MPI_SENDRECV halo callback (CUDA-aware MPI).
!
! This image has no Fortran compiler, so the file is syntax-simple F2003 + cpp and is not built here; the same C entry
! points are exercised through ctypes by tests/ (tests/test_abi.py checks that every bind(C) name below is declared in
! include/flutas_b200.h and exported by the library).
!
#if defined(B200_REPLACE_MOD_BOUND)
#define B200_MOD_BOUND mod_bound
#else
#define B200_MOD_BOUND mod_bound_b200
#endif
#if defined(B200_REPLACE_MOD_CHKDT)
#define B200_MOD_CHKDT mod_chkdt
#else
#define B200_MOD_CHKDT mod_chkdt_b200
#endif
module mod_flutas_b200
  use, intrinsic :: iso_c_binding
  implicit none
  interface
    integer(c_int) function flutas_b200_init(device,rank,nranks) bind(C,name='flutas_b200_init')
      import; integer(c_int), value :: device,rank,nranks
    end function
    integer(c_int) function flutas_b200_synchronize() bind(C,name='flutas_b200_synchronize')
      import
    end function
    type(c_ptr) function flutas_b200_alloc_managed(bytes) bind(C,name='flutas_b200_alloc_managed')
      import; integer(c_size_t), value :: bytes
    end function
    subroutine flutas_b200_free(ptr) bind(C,name='flutas_b200_free')
      import; type(c_ptr), value :: ptr
    end subroutine
    integer(c_int) function flutas_b200_fftini(n_x,n_y,bcxy,c_or_f,arrplan,normfft) bind(C,name='flutas_b200_fftini')
      import; integer(c_int), intent(in) :: n_x(3),n_y(3)
      character(kind=c_char), intent(in) :: bcxy(4),c_or_f(2)
      type(c_ptr), intent(out) :: arrplan(4); real(c_double), intent(out) :: normfft
    end function
    integer(c_int) function flutas_b200_fftend(arrplan) bind(C,name='flutas_b200_fftend')
      import; type(c_ptr), intent(inout) :: arrplan(4)
    end function
    integer(c_int) function flutas_b200_fft(plan,n,arr) bind(C,name='flutas_b200_fft')
      import; type(c_ptr), value :: plan,arr; integer(c_int), intent(in) :: n(3)
    end function
    integer(c_int) function flutas_b200_solver(n,arrplan,normfft,lambdaxy,a,b,c,bcz,c_or_f,p) bind(C,name='flutas_b200_solver')
      import; integer(c_int), intent(in) :: n(3); type(c_ptr), intent(in) :: arrplan(4)
      real(c_double), value :: normfft; type(c_ptr), value :: lambdaxy,a,b,c,p
      character(kind=c_char), intent(in) :: bcz(2),c_or_f(3)
    end function
    integer(c_int) function flutas_b200_solver_slab(n,arrplan,normfft,lambdaxy_g,a,b,c,bcz,c_or_f,p) &
                            bind(C,name='flutas_b200_solver_slab')
      import; integer(c_int), intent(in) :: n(3); type(c_ptr), intent(in) :: arrplan(4)
      real(c_double), value :: normfft; type(c_ptr), value :: lambdaxy_g,a,b,c,p
      character(kind=c_char), intent(in) :: bcz(2),c_or_f(3)
    end function
    integer(c_size_t) function flutas_b200_p2p_handle_bytes() bind(C,name='flutas_b200_p2p_handle_bytes')
      import
    end function
    integer(c_int) function flutas_b200_p2p_export(arrplan,n_local,blob) bind(C,name='flutas_b200_p2p_export')
      import; type(c_ptr), intent(in) :: arrplan(4); integer(c_int), intent(in) :: n_local(3); type(c_ptr), value :: blob
    end function
    integer(c_int) function flutas_b200_p2p_attach(arrplan,blobs) bind(C,name='flutas_b200_p2p_attach')
      import; type(c_ptr), intent(in) :: arrplan(4); type(c_ptr), value :: blobs
    end function
    integer(c_int) function flutas_b200_set_halo_exchange(fn,ctx) bind(C,name='flutas_b200_set_halo_exchange')
      import; type(c_funptr), value :: fn; type(c_ptr), value :: ctx
    end function
    integer(c_int) function flutas_b200_fillps(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzfi,dti,rho0,u,v,w,p) &
                            bind(C,name='flutas_b200_fillps')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi,dti,rho0
      type(c_ptr), value :: dzfi,u,v,w,p
    end function
    integer(c_int) function flutas_b200_updt_rhs_b(nx,ny,nz,cbc,rhsbx,rhsby,rhsbz,p) bind(C,name='flutas_b200_updt_rhs_b')
      import; integer(c_int), value :: nx,ny,nz; character(kind=c_char), intent(in) :: cbc(6)
      type(c_ptr), value :: rhsbx,rhsby,rhsbz,p
    end function
    integer(c_int) function flutas_b200_correc(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzci,dt,rho0,p,u,v,w,rho) &
                            bind(C,name='flutas_b200_correc')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi,dt,rho0
      type(c_ptr), value :: dzci,p,u,v,w,rho
    end function
    integer(c_int) function flutas_b200_chkdiv(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzfi,u,v,w,divtot,divmax) &
                            bind(C,name='flutas_b200_chkdiv')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi
      type(c_ptr), value :: dzfi,u,v,w; real(c_double), intent(out) :: divtot,divmax
    end function
    integer(c_int) function flutas_b200_chkdt(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,dzfi,u,v,w,dti) &
                            bind(C,name='flutas_b200_chkdt')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi
      type(c_ptr), value :: dzci,dzfi,u,v,w; real(c_double), intent(out) :: dti
    end function
    integer(c_int) function flutas_b200_bounduvw(cbc,n,bc,nh_d,nh_u,isoutflow,dl,dzc,dzf,u,v,w) &
                            bind(C,name='flutas_b200_bounduvw')
      import; character(kind=c_char), intent(in) :: cbc(18); integer(c_int), intent(in) :: n(3),isoutflow(6)
      real(c_double), intent(in) :: bc(18),dl(3); integer(c_int), value :: nh_d,nh_u
      type(c_ptr), value :: dzc,dzf,u,v,w
    end function
    integer(c_int) function flutas_b200_pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,pold,u,v,w) &
                            bind(C,name='flutas_b200_pres_sp_src')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: f_t12,dxi,dyi,dzi,rho0i
      type(c_ptr), value :: dzci,pold,u,v,w
    end function
    integer(c_int) function flutas_b200_pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,f_t12,f_t12_o,p,pold,rho,u,v,w) &
                            bind(C,name='flutas_b200_pres_tw_src')
      import; integer(c_int), value :: nx,ny,nz,nh_d,nh_u; real(c_double), value :: dxi,dyi,dzi,rho0i,f_t12,f_t12_o
      type(c_ptr), value :: dzci,p,pold,rho,u,v,w
    end function
    integer(c_int) function flutas_b200_load(io,filename,ng,n,start,nh,fld) bind(C,name='flutas_b200_load')
      import; character(kind=c_char), value :: io; character(kind=c_char), intent(in) :: filename(*)
      integer(c_int), intent(in) :: ng(3),n(3),start(3); integer(c_int), value :: nh; type(c_ptr), value :: fld
    end function
    integer(c_int) function flutas_b200_pold_update(nx,ny,nz,mode,p,pold) bind(C,name='flutas_b200_pold_update')
      import; integer(c_int), value :: nx,ny,nz,mode; type(c_ptr), value :: p,pold
    end function
    integer(c_int) function flutas_b200_boundp(cbc,n,bc,nh_d,nh_p,dl,dzc,dzf,p) bind(C,name='flutas_b200_boundp')
      import; character(kind=c_char), intent(in) :: cbc(6); integer(c_int), intent(in) :: n(3)
      real(c_double), intent(in) :: bc(6),dl(3); integer(c_int), value :: nh_d,nh_p
      type(c_ptr), value :: dzc,dzf,p
    end function
    function flutas_b200_last_error() bind(C,name='flutas_b200_last_error') result(msg)
      import; type(c_ptr) :: msg
    end function
  end interface
  logical, save :: b200_ready = .false.
contains
  subroutine b200_check(istat,where)          ! the reference prints and stops on errors (src/fft.f90:879-883)
    integer(c_int), intent(in) :: istat
    character(len=*), intent(in) :: where
    character(kind=c_char), pointer :: msg(:)
    integer :: q
    if(istat.ne.0) then
      call c_f_pointer(flutas_b200_last_error(),msg,(/512/))
      q = 1
      do while(q.lt.512.and.msg(q).ne.c_null_char)
        q = q + 1
      enddo
      print*, 'flutas_b200 error in ', where, ': ', msg(1:q-1)
      error stop 1
    endif
  end subroutine b200_check
  !
  ! GPU binding of initmpi (src/initmpi.f90:59-63: device = rank local to the node) + the slab decomposition
  ! dims_in = (1,nranks); called by fftini, i.e. from initsolver, before any kernel runs
  !
  subroutine b200_setup()
    use mpi
    integer :: ierr,myid,nproc,local_comm,mydev
    if(b200_ready) return
    call MPI_COMM_RANK(MPI_COMM_WORLD,myid ,ierr)
    call MPI_COMM_SIZE(MPI_COMM_WORLD,nproc,ierr)
    call MPI_COMM_SPLIT_TYPE(MPI_COMM_WORLD,MPI_COMM_TYPE_SHARED,0,MPI_INFO_NULL,local_comm,ierr)
    call MPI_COMM_RANK(local_comm,mydev,ierr)
    call b200_check(flutas_b200_init(int(mydev,c_int),int(myid,c_int),int(nproc,c_int)),'init')
    b200_ready = .true.
  end subroutine b200_setup
  !
  ! fields in managed memory, Fortran bounds (lo:hi)^3 -- replaces `allocate(p(0:n(1)+1,0:n(2)+1,0:n(3)+1))` etc. in the main
  !
  subroutine b200_alloc_field(fld,lo,hi)
    real(c_double), pointer, intent(out) :: fld(:,:,:)
    integer, intent(in) :: lo(3),hi(3)
    real(c_double), pointer :: flat(:,:,:)
    type(c_ptr) :: raw
    integer(c_size_t) :: bytes
    call b200_setup()
    bytes = 8_c_size_t*int(hi(1)-lo(1)+1,c_size_t)*int(hi(2)-lo(2)+1,c_size_t)*int(hi(3)-lo(3)+1,c_size_t)
    raw = flutas_b200_alloc_managed(bytes)
    if(.not.c_associated(raw)) call b200_check(1_c_int,'alloc_field')
    call c_f_pointer(raw,flat,(/hi(1)-lo(1)+1,hi(2)-lo(2)+1,hi(3)-lo(3)+1/))
    fld(lo(1):,lo(2):,lo(3):) => flat
    fld = 0._c_double
  end subroutine b200_alloc_field
  !
  ! z-halo exchange of boundp / bounduvw on several ranks: the MPI_SENDRECV pair of updthalo (src/bound.f90:1098-1103) on
  ! device pointers (CUDA-aware MPI); lo / hi = -1 -> MPI_PROC_NULL.  Registered by b200_slab_attach.
  !
  function b200_halo_cb(ctx,send_lo,send_hi,recv_lo,recv_hi,count,lo,hi,stream) bind(C) result(rc)
    use mpi
    type(c_ptr), value :: ctx,send_lo,send_hi,recv_lo,recv_hi,stream
    integer(c_size_t), value :: count
    integer(c_int), value :: lo,hi
    integer(c_int) :: rc
    real(c_double), pointer :: slo(:),shi(:),rlo(:),rhi(:)
    integer :: ierr,nlo,nhi,status(MPI_STATUS_SIZE)
    rc = flutas_b200_synchronize()                       ! the planes to send are produced on the library stream
    nlo = MPI_PROC_NULL; if(lo.ge.0) nlo = lo
    nhi = MPI_PROC_NULL; if(hi.ge.0) nhi = hi
    call c_f_pointer(send_lo,slo,(/count/)); call c_f_pointer(send_hi,shi,(/count/))
    call c_f_pointer(recv_lo,rlo,(/count/)); call c_f_pointer(recv_hi,rhi,(/count/))
    call MPI_SENDRECV(slo,int(count),MPI_DOUBLE_PRECISION,nlo,0,rhi,int(count),MPI_DOUBLE_PRECISION,nhi,0, &
                      MPI_COMM_WORLD,status,ierr)
    call MPI_SENDRECV(shi,int(count),MPI_DOUBLE_PRECISION,nhi,1,rlo,int(count),MPI_DOUBLE_PRECISION,nlo,1, &
                      MPI_COMM_WORLD,status,ierr)
  end function b200_halo_cb
end module mod_flutas_b200
!
module mod_fft                                ! same public names as src/fft.f90:17
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: fftini,fftend,fft
contains
  subroutine fft(plan,arr)                    ! src/fft.f90:181-193: one unnormalised in-place transform, FFTW element order
    type(C_PTR), intent(in   )                           :: plan
    real(rp)   , intent(inout), dimension(:,:,:), target :: arr
    call b200_check(flutas_b200_fft(plan,int(shape(arr),c_int),c_loc(arr)),'fft')
  end subroutine fft
  subroutine fftini(n_x,n_y,bcxy,c_or_f,arrplan,normfft)
    integer         , intent(in ), dimension(3)     :: n_x,n_y
    character(len=1), intent(in ), dimension(0:1,2) :: bcxy
    character(len=1), intent(in ), dimension(2)     :: c_or_f
    type(C_PTR)     , intent(out), dimension(2,2)   :: arrplan
    real(rp)        , intent(out)                   :: normfft
    type(C_PTR) :: plans(4)
    character(kind=c_char) :: cb(4),cf(2)
    call b200_setup()
    cb = (/bcxy(0,1),bcxy(1,1),bcxy(0,2),bcxy(1,2)/)
    cf = (/c_or_f(1),c_or_f(2)/)
    call b200_check(flutas_b200_fftini(int(n_x,c_int),int(n_y,c_int),cb,cf,plans,normfft),'fftini')
    arrplan = reshape(plans,(/2,2/))           ! (1,1) fwd-x (2,1) bwd-x (1,2) fwd-y (2,2) bwd-y, as fft.f90:151-154
  end subroutine fftini
  subroutine fftend(arrplan)
    type(C_PTR), intent(in), dimension(2,2) :: arrplan
    type(C_PTR) :: plans(4)
    plans = reshape(arrplan,(/4/))
    call b200_check(flutas_b200_fftend(plans),'fftend')
  end subroutine fftend
end module mod_fft
!
! common back end of solver_cpu / solver_gpu: one rank -> flutas_b200_solver; dims_in = (1,N) -> flutas_b200_solver_slab
! with the peers' exchange buffers attached and the eigenvalue windows all-gathered on the first call
!
module mod_solver_b200
  use, intrinsic :: iso_c_binding
  use mpi
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: b200_solve
  real(rp), allocatable, target, save :: lambdaxy_g(:,:)
  logical, save :: attached = .false.
contains
  subroutine b200_solve(n,arrplan,normfft,lambdaxy,a,b,c,bcz,c_or_f,p)
    integer         , intent(in   ), dimension(3)             :: n          ! local x-pencil interior size
    type(C_PTR)     , intent(in   ), dimension(2,2)           :: arrplan
    real(rp)        , intent(in   )                           :: normfft
    real(rp)        , intent(in   ), dimension(:,:), target   :: lambdaxy   ! z-pencil window (ng1, ng2/N), CPU order
    real(rp)        , intent(in   ), dimension(:)  , target   :: a,b,c
    character(len=1), intent(in   ), dimension(0:1)           :: bcz
    character(len=1), intent(in   ), dimension(3)             :: c_or_f
    real(rp)        , intent(inout), dimension(0:,0:,0:), target :: p
    type(C_PTR) :: plans(4)
    character(kind=c_char) :: bz(2),cf(3)
    character(kind=c_char), allocatable, target :: blob(:),blobs(:)
    integer :: nproc,ierr,nb
    plans = reshape(arrplan,(/4/))
    bz = (/bcz(0),bcz(1)/); cf = c_or_f
    call MPI_COMM_SIZE(MPI_COMM_WORLD,nproc,ierr)
    if(nproc.eq.1) then
      call b200_check(flutas_b200_solver(int(n,c_int),plans,normfft,c_loc(lambdaxy),c_loc(a),c_loc(b),c_loc(c), &
                                         bz,cf,c_loc(p)),'solver')
      return
    endif
    if(.not.attached) then
      ! (a) CUDA IPC handles of every rank's exchange memory, rank order
      nb = int(flutas_b200_p2p_handle_bytes())
      allocate(blob(nb),blobs(nb*nproc))
      call b200_check(flutas_b200_p2p_export(plans,int(n,c_int),c_loc(blob)),'p2p_export')
      call MPI_ALLGATHER(blob,nb,MPI_BYTE,blobs,nb,MPI_BYTE,MPI_COMM_WORLD,ierr)
      call b200_check(flutas_b200_p2p_attach(plans,c_loc(blobs)),'p2p_attach')
      deallocate(blob,blobs)
      ! (b) eigenvalues: initsolver gives every rank the window lambdaxy(ng1, ng2/N) of its z-pencil (initsolver.f90:87-93);
      !     the x-split z stage needs all of y -> concatenate the windows along y (contiguous in column-major order)
      allocate(lambdaxy_g(size(lambdaxy,1),size(lambdaxy,2)*nproc))
      call MPI_ALLGATHER(lambdaxy,size(lambdaxy),MPI_REAL_RP,lambdaxy_g,size(lambdaxy),MPI_REAL_RP,MPI_COMM_WORLD,ierr)
      ! (c) z-halo planes of boundp / bounduvw
      call b200_check(flutas_b200_set_halo_exchange(c_funloc(b200_halo_cb),c_null_ptr),'set_halo_exchange')
      attached = .true.
    endif
    call b200_check(flutas_b200_solver_slab(int(n,c_int),plans,normfft,c_loc(lambdaxy_g),c_loc(a),c_loc(b),c_loc(c), &
                                            bz,cf,c_loc(p)),'solver_slab')
  end subroutine b200_solve
end module mod_solver_b200
!
module mod_solver_cpu                         ! same signature as src/solver_cpu.f90:20-31
  use, intrinsic :: iso_c_binding
  use mod_solver_b200
  use mod_common_mpi, only: n_z
  use mod_types
  implicit none
  private
  public :: solver_cpu
contains
  subroutine solver_cpu(n,arrplan,normfft,lambdaxy,a,b,c,bcz,c_or_f,p)
    integer         , intent(in   ), dimension(3)                       :: n
    type(C_PTR)     , intent(in   ), dimension(2,2)                     :: arrplan
    real(rp)        , intent(in   )                                     :: normfft
    real(rp)        , intent(in   ), dimension(n_z(1),n_z(2)), target   :: lambdaxy
    real(rp)        , intent(in   ), dimension(n_z(3))       , target   :: a,b,c
    character(len=1), intent(in   ), dimension(0:1)                     :: bcz
    character(len=1), intent(in   ), dimension(3)                       :: c_or_f
    real(rp)        , intent(inout), dimension(0:,0:,0:)     , target   :: p
    call b200_solve(n,arrplan,normfft,lambdaxy,a,b,c,bcz,c_or_f,p)
  end subroutine solver_cpu
end module mod_solver_cpu
!
module mod_solver_gpu                         ! same signature as src/solver_gpu.f90:31-47 (n = n_z, bc(0:1,3))
  use, intrinsic :: iso_c_binding
  use mod_solver_b200
  use mod_common_mpi, only: n_x
  use mod_types
  implicit none
  private
  public :: solver_gpu
contains
  subroutine solver_gpu(n,dims,arrplan,normfft,lambdaxy,a,b,c,bc,c_or_f,p)
    integer         , intent(in   ), dimension(3)             :: n,dims
    type(C_PTR)     , intent(in   ), dimension(2,2)           :: arrplan
    real(rp)        , intent(in   )                           :: normfft
    real(rp)        , intent(in   ), dimension(:,:), target   :: lambdaxy
    real(rp)        , intent(in   ), dimension(:)  , target   :: a,b,c
    character(len=1), intent(in   ), dimension(0:1,3)         :: bc
    character(len=1), intent(in   ), dimension(3)             :: c_or_f
    real(rp)        , intent(inout), dimension(0:,0:,0:), target :: p
    ! the reference passes n = n_z; the x-pencil size the library wants is mod_common_mpi's n_x.  lambdaxy must be in the
    ! CPU (FFTW half-complex) order of initsolver.f90:136-139, i.e. initsolver compiled without _OPENACC.
    call b200_solve(n_x,arrplan,normfft,lambdaxy,a,b,c,bc(:,3),c_or_f,p)
  end subroutine solver_gpu
end module mod_solver_gpu
!
module mod_fillps                             ! same signature as src/fillps.f90:16-26
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: fillps
contains
  subroutine fillps(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzfi,dti,rho0,u,v,w,p)
    integer , intent(in )                                     :: nx,ny,nz,nh_d,nh_u
    real(rp), intent(in )                                     :: dxi,dyi,dzi,dti,rho0
    real(rp), intent(in ), dimension(1-nh_d:), target         :: dzfi
    real(rp), intent(in ), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(out), dimension(0:,0:,0:), target        :: p
    call b200_check(flutas_b200_fillps(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,c_loc(dzfi),dti,rho0, &
                                       c_loc(u),c_loc(v),c_loc(w),c_loc(p)),'fillps')
  end subroutine fillps
end module mod_fillps
!
module mod_correc                             ! same signature as src/correc.f90:16-29
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_types
  implicit none
  private
  public :: correc
contains
  subroutine correc(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,dzci,dt,rho0,p,u,v,w,rho)
    integer , intent(in   )                                     :: nx,ny,nz,nh_d,nh_u
    real(rp), intent(in   )                                     :: dxi,dyi,dzi,dt,rho0
    real(rp), intent(in   ), dimension(1-nh_d:), target         :: dzci
    real(rp), intent(in   ), dimension(0:,0:,0:), target        :: p
    real(rp), intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(in   ), dimension(0:,0:,0:)                :: rho      ! (0,0,0)-sized dummy: never touched
    call b200_check(flutas_b200_correc(nx,ny,nz,nh_d,nh_u,dxi,dyi,dzi,c_loc(dzci),dt,rho0, &
                                       c_loc(p),c_loc(u),c_loc(v),c_loc(w),c_null_ptr),'correc')
  end subroutine correc
end module mod_correc
!
module mod_chkdiv                             ! same signature as src/chkdiv.f90:18-29
  use, intrinsic :: iso_c_binding
  use mpi
  use mod_flutas_b200
  use mod_common_mpi, only: myid,ierr
  use mod_types
  implicit none
  private
  public :: chkdiv
contains
  subroutine chkdiv(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzfi,u,v,w,divtot,divmax)
    integer , intent(in )                                     :: nx,ny,nz,nh_d,nh_u
    real(rp), intent(in )                                     :: dxi,dyi,dzi
    real(rp), intent(in ), dimension(1-nh_d:), target         :: dzfi
    real(rp), intent(in ), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(out)                                     :: divtot,divmax
    call b200_check(flutas_b200_chkdiv(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzfi),c_loc(u),c_loc(v),c_loc(w), &
                                       divtot,divmax),'chkdiv')
    call mpi_allreduce(MPI_IN_PLACE,divtot,1,MPI_REAL_RP,MPI_SUM,MPI_COMM_WORLD,ierr)   ! as chkdiv.f90:64-65
    call mpi_allreduce(MPI_IN_PLACE,divmax,1,MPI_REAL_RP,MPI_MAX,MPI_COMM_WORLD,ierr)
    if(myid.eq.0) print*, 'Total divergence = ', divtot, '| Maximum divergence = ', divmax
  end subroutine chkdiv
end module mod_chkdiv
!
! chkdt_sp, same signature as src/chkdt.f90:117: the field reduction runs on the device, the all-reduce and the scalar
! formulas (:183-196) are the reference's own lines
!
module B200_MOD_CHKDT
  use, intrinsic :: iso_c_binding
  use mpi
  use mod_flutas_b200
  use mod_common_mpi, only: ierr
  use mod_param     , only: rho_sp,mu_sp,cfl_c,cfl_d,gacc_x,gacc_y,gacc_z
  use mod_types
  implicit none
  private
  public :: chkdt_sp
contains
  subroutine chkdt_sp(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,dzfi,u,v,w,dtmax)
    integer , intent(in )                                     :: nx,ny,nz
    real(rp), intent(in )                                     :: dxi,dyi,dzi
    integer , intent(in )                                     :: nh_d,nh_u
    real(rp), intent(in ), dimension(1-nh_d:), target         :: dzci,dzfi
    real(rp), intent(in ), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    real(rp), intent(out)                                     :: dtmax
    real(rp) :: dti,dtiv,dtig,dlmin,dlmini
    call b200_check(flutas_b200_chkdt(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzci),c_loc(dzfi),c_loc(u),c_loc(v),c_loc(w), &
                                      dti),'chkdt')
    call mpi_allreduce(MPI_IN_PLACE,dti,1,MPI_REAL_RP,MPI_MAX,MPI_COMM_WORLD,ierr)       ! chkdt.f90:183
    if(dti.eq.0._rp) dti = 1._rp
    dlmin  = min(1._rp/dxi,1._rp/dyi,1._rp/dzi)
    dlmin  = min(dlmin,minval(1._rp/dzfi(:)))
    dlmini = dlmin**(-1)
    dtiv   = (mu_sp/rho_sp)*dlmini**2
    dtig   = sqrt( max(abs(gacc_x),abs(gacc_y),abs(gacc_z))*dlmini )
    dtmax  = min(cfl_c/dti,cfl_d/dtiv,1._rp/dtig)
  end subroutine chkdt_sp
end module B200_MOD_CHKDT
!
! boundp (pressure halo, nh_p = 1), bounduvw (any halo width) and updt_rhs_b: same argument lists as src/bound.f90:146,17,829;
! `halo` (MPI datatypes) is unused.  cbc / bc are passed in Fortran storage order, which is what the C side expects.
!
module B200_MOD_BOUND
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  implicit none
  private
  public :: boundp,bounduvw,updt_rhs_b
  contains
  subroutine boundp(cbc,n,bc,nh_d,nh_p,halo,dl,dzc,dzf,p)
    character(len=1), intent(in   ), dimension(0:1,3)                   :: cbc
    integer         , intent(in   ), dimension(3)                       :: n
    real(c_double)  , intent(in   ), dimension(0:1,3)                   :: bc
    integer         , intent(in   )                                     :: nh_d,nh_p
    integer         , intent(in   ), dimension(3)                       :: halo
    real(c_double)  , intent(in   ), dimension(3)                       :: dl
    real(c_double)  , intent(in   ), dimension(1-nh_d:), target         :: dzc,dzf
    real(c_double)  , intent(inout), dimension(1-nh_p:,1-nh_p:,1-nh_p:), target :: p
    character(kind=c_char) :: cc(6)
    real(c_double) :: bv(6)
    cc = reshape(cbc,(/6/))
    bv = reshape(bc ,(/6/))
    call b200_check(flutas_b200_boundp(cc,int(n,c_int),bv,int(nh_d,c_int),int(nh_p,c_int),dl, &
                                       c_loc(dzc),c_loc(dzf),c_loc(p)),'boundp')
  end subroutine boundp
  subroutine bounduvw(cbc,n,bc,nh_d,nh_u,halo,isoutflow,dl,dzc,dzf,u,v,w)
    character(len=1), intent(in   ), dimension(0:1,3,3)                 :: cbc
    integer         , intent(in   ), dimension(3)                       :: n
    real(c_double)  , intent(in   ), dimension(0:1,3,3)                 :: bc
    integer         , intent(in   )                                     :: nh_d,nh_u
    integer         , intent(in   ), dimension(3)                       :: halo
    logical         , intent(in   ), dimension(0:1,3)                   :: isoutflow
    real(c_double)  , intent(in   ), dimension(3)                       :: dl
    real(c_double)  , intent(in   ), dimension(1-nh_d:), target         :: dzc,dzf
    real(c_double)  , intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    character(kind=c_char) :: cc(18)
    real(c_double) :: bv(18)
    integer(c_int) :: io(6)
    cc = reshape(cbc,(/18/))
    bv = reshape(bc ,(/18/))
    io = merge(1_c_int,0_c_int,reshape(isoutflow,(/6/)))
    call b200_check(flutas_b200_bounduvw(cc,int(n,c_int),bv,int(nh_d,c_int),int(nh_u,c_int),io,dl, &
                                         c_loc(dzc),c_loc(dzf),c_loc(u),c_loc(v),c_loc(w)),'bounduvw')
  end subroutine bounduvw
  subroutine updt_rhs_b(nx,ny,nz,c_or_f,cbc,nh_p,rhsbx,rhsby,rhsbz,p)
    integer         , intent(in   )                                     :: nx,ny,nz
    character       , intent(in   ), dimension(3)                       :: c_or_f   ! 'c','c','c' in every app
    character(len=1), intent(in   ), dimension(0:1,3)                   :: cbc
    integer         , intent(in   )                                     :: nh_p
    real(c_double)  , intent(in   ), dimension(      :,      :,     0:), target :: rhsbx,rhsby,rhsbz
    real(c_double)  , intent(inout), dimension(1-nh_p:,1-nh_p:,1-nh_p:), target :: p
    character(kind=c_char) :: cc(6)
    if(any(c_or_f.ne.'c').or.nh_p.ne.1) call b200_check(2_c_int,'updt_rhs_b (cell-centred pressure with nh_p = 1 only)')
    cc = reshape(cbc,(/6/))
    call b200_check(flutas_b200_updt_rhs_b(nx,ny,nz,cc,c_loc(rhsbx),c_loc(rhsby),c_loc(rhsbz),c_loc(p)),'updt_rhs_b')
  end subroutine updt_rhs_b
end module B200_MOD_BOUND

!
! pressure-gradient source terms of the predictor, same signatures as src/source.f90:247,311 (constant-coefficient
! Poisson build), and the two pressure bookkeeping loops of main__single_phase.f90:693-699 / :734-740 as
!   call pold_update(n,0,p,pold)   ! pold = p        call pold_update(n,1,p,pold)   ! p = pold + p
!
module mod_source_b200
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  implicit none
  private
  public :: pres_sp_src,pres_tw_src,pold_update
  contains
  subroutine pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,pold,u,v,w)
    integer       , intent(in   )                                     :: nx,ny,nz,nh_d,nh_u
    real(c_double), intent(in   )                                     :: f_t12,dxi,dyi,dzi,rho0i
    real(c_double), intent(in   ), dimension(1-nh_d:), target         :: dzci
    real(c_double), intent(in   ), dimension(0:,0:,0:), target        :: pold
    real(c_double), intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    call b200_check(flutas_b200_pres_sp_src(nx,ny,nz,f_t12,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzci),rho0i, &
                                            c_loc(pold),c_loc(u),c_loc(v),c_loc(w)),'pres_sp_src')
  end subroutine pres_sp_src
  subroutine pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,dzci,rho0i,f_t12,f_t12_o,p,pold,rho,u,v,w)
    integer       , intent(in   )                                     :: nx,ny,nz,nh_d,nh_u
    real(c_double), intent(in   )                                     :: dxi,dyi,dzi,rho0i,f_t12,f_t12_o
    real(c_double), intent(in   ), dimension(1-nh_d:), target         :: dzci
    real(c_double), intent(in   ), dimension(0:,0:,0:), target        :: p,pold,rho
    real(c_double), intent(inout), dimension(1-nh_u:,1-nh_u:,1-nh_u:), target :: u,v,w
    call b200_check(flutas_b200_pres_tw_src(nx,ny,nz,dxi,dyi,dzi,nh_d,nh_u,c_loc(dzci),rho0i,f_t12,f_t12_o, &
                                            c_loc(p),c_loc(pold),c_loc(rho),c_loc(u),c_loc(v),c_loc(w)),'pres_tw_src')
  end subroutine pres_tw_src
  subroutine pold_update(n,mode,p,pold)
    integer       , intent(in   ), dimension(3)                 :: n
    integer       , intent(in   )                               :: mode
    real(c_double), intent(inout), dimension(0:,0:,0:), target  :: p,pold
    call b200_check(flutas_b200_pold_update(n(1),n(2),n(3),mode,c_loc(p),c_loc(pold)),'pold_update')
  end subroutine pold_update
end module mod_source_b200

!
! restart files: same signature as src/load.f90:21 (decomposition from 2DECOMP's xstart, global size from mod_param)
!
module mod_load_b200
  use, intrinsic :: iso_c_binding
  use mod_flutas_b200
  use mod_param , only: ng
  use decomp_2d , only: xstart
  implicit none
  private
  public :: load
  contains
  subroutine load(io,filename,n,fld)
    character(len=1), intent(in   )                                    :: io
    character(len=*), intent(in   )                                    :: filename
    integer         , intent(in   ), dimension(3)                      :: n
    real(c_double)  , intent(inout), dimension(n(1),n(2),n(3)), target :: fld
    call b200_check(flutas_b200_load(io,trim(filename)//c_null_char,int(ng,c_int),int(n,c_int),int(xstart-1,c_int), &
                                     0_c_int,c_loc(fld)),'load')
  end subroutine load
end module mod_load_b200
