"""A MIPS test program touching every instruction class tests/cpu_gen.py interprets (straight-line except for the branch
and jump forms, each of which skips one instruction after its delay slot)."""
from cpu_gen import rtype, itype, jtype

ENTRY, DATA = 0x1000, 0x2000
SYNC = rtype(0b001111)


KDATA, KOUT, WDATA, HDATA, IMAGE_ID = 0x2100, 0x2200, 0x2300, 0x2400, 0x2500
KECCAK_LEN = 140


def build(with_syscalls=False, sha_blocks=0):
    p = []

    def emit(w):
        p.append(w & 0xFFFFFFFF)
        return ENTRY + 4 * (len(p) - 1)
    # constants
    emit(itype(0b001111, 0, 1, 0x1234))                 # lui  $1, 0x1234
    emit(itype(0b001101, 1, 1, 0x5678))                 # ori  $1, $1, 0x5678
    emit(itype(0b001001, 0, 2, 0xFFFB))                 # addiu $2, $0, -5
    emit(itype(0b001000, 0, 3, 100))                    # addi $3, $0, 100
    emit(itype(0b001111, 0, 27, 0x8000))                # lui  $27, 0x8000 (sign-extended immediate)
    # register arithmetic
    for func, rd, rs, rt in ((0b100000, 4, 1, 3), (0b100001, 5, 1, 2), (0b100010, 6, 1, 3), (0b100011, 7, 3, 1), (0b101010, 8, 2, 3),
                             (0b101011, 9, 2, 3), (0b101010, 8, 27, 3), (0b100000, 4, 27, 27)):
        emit(rtype(func, rs, rt, rd))
    emit(itype(0b001010, 2, 10, 7))                     # slti
    emit(itype(0b001011, 3, 11, 0xFFFF))                # sltiu
    emit(itype(0b001000, 1, 12, 0x8001))                # addi with negative immediate
    for func, rd in ((0b100100, 12), (0b100101, 13), (0b100110, 14), (0b100111, 15)):
        emit(rtype(func, 1, 2, rd))                     # and or xor nor
    emit(itype(0b001100, 1, 12, 0xFF0F))                # andi
    emit(itype(0b001110, 1, 13, 0xFFFF))                # xori
    emit(rtype(0b000010, 1, 3, 16, opcode=0b011100))    # mul
    emit(rtype(0b011000, 1, 2))                         # mult
    emit(rtype(0b010000, rd=17))                        # mfhi
    emit(rtype(0b010010, rd=18))                        # mflo
    emit(rtype(0b011001, 1, 2))                         # multu
    emit(rtype(0b011010, 2, 3))                         # div  (-5 / 100)
    emit(rtype(0b011010, 1, 2))                         # div  (positive / negative)
    emit(rtype(0b011011, 1, 3))                         # divu
    emit(rtype(0b010001, rs=3))                         # mthi
    emit(rtype(0b010011, rs=1))                         # mtlo
    emit(rtype(0b000001, 1, 3, opcode=0b011100))        # maddu
    emit(rtype(0b000001, 2, 2, opcode=0b011100))        # maddu, large product
    # shifts
    emit(rtype(0b000000, 0, 1, 19, 5))                  # sll
    emit(rtype(0b000010, 0, 1, 20, 7))                  # srl
    emit(rtype(0b000011, 0, 2, 21, 3))                  # sra
    emit(rtype(0b000011, 0, 1, 21, 0))                  # sra by 0
    emit(itype(0b001001, 0, 22, 9))                     # addiu $22, $0, 9
    emit(rtype(0b000100, 22, 1, 23))                    # sllv
    emit(rtype(0b000110, 22, 1, 24))                    # srlv
    emit(rtype(0b000111, 22, 2, 25))                    # srav
    emit(rtype(0b000010, 1, 1, 26, 8))                  # rotr $26, $1, 8
    # conditional moves, counts, bit fields
    emit(rtype(0b001010, 1, 0, 4))                      # movz (moves)
    emit(rtype(0b001010, 1, 3, 4))                      # movz (does not)
    emit(rtype(0b001011, 2, 3, 5))                      # movn (moves)
    emit(rtype(0b001011, 2, 0, 5))                      # movn (does not)
    emit(rtype(0b100000, 1, 0, 6, opcode=0b011100))     # clz
    emit(rtype(0b100001, 2, 0, 7, opcode=0b011100))     # clo
    emit(rtype(0b100000, 0, 0, 6, opcode=0b011100))     # clz of 0
    emit(rtype(0b000000, 1, 8, 7, 4, opcode=0b011111))  # ext $8, $1, 4, 8
    emit(rtype(0b000100, 2, 1, 11, 4, opcode=0b011111))  # ins $1, $2, 4, 8
    emit(rtype(0b100000, 0, 2, 9, 0b010000, opcode=0b011111))   # seb
    emit(rtype(0b100000, 0, 1, 10, 0b011000, opcode=0b011111))  # seh
    emit(rtype(0b100000, 0, 1, 11, 0b000010, opcode=0b011111))  # wsbh
    emit(rtype(0b111011, 0, 12, 0, opcode=0b011111))    # rdhwr $12, $0
    emit(rtype(0b111011, 0, 12, 29, opcode=0b011111))   # rdhwr $12, $29
    emit(rtype(0b111011, 0, 12, 5, opcode=0b011111))    # rdhwr $12, $5
    emit(rtype(0b110100, 1, 3))                         # teq (no trap)
    emit(SYNC)
    emit(itype(0b110011, 0, 0, 0))                      # pref = nop
    # memory
    emit(itype(0b001101, 0, 28, DATA))                  # ori $28, $0, DATA
    for opcode, off in ((0b100011, 0), (0b100000, 1), (0b100000, 6), (0b100100, 3), (0b100001, 2), (0b100001, 4), (0b100101, 6),
                        (0b100010, 1), (0b100010, 7), (0b100110, 2), (0b100110, 4), (0b110000, 8)):
        emit(itype(opcode, 28, 14, off))                # lw lb lb lbu lh lh lhu lwl lwl lwr lwr ll
    for opcode, off in ((0b101011, 12), (0b101000, 13), (0b101000, 18), (0b101001, 14), (0b101001, 16), (0b101010, 21), (0b101010, 23),
                        (0b101110, 25), (0b101110, 26), (0b111000, 28), (0b111101, 32)):
        emit(itype(opcode, 28, 1, off))                 # sw sb sb sh sh swl swl swr swr sc sdc1
    emit(itype(0b100011, 28, 15, 12))                   # lw back what sw wrote
    emit(itype(0b100011, 28, 15, 0xFFFC))               # lw with a negative offset (DATA - 4)
    # control flow: each taken branch/jump skips the instruction after its delay slot
    for opcode, rs, rt in ((0x04, 3, 3), (0x04, 1, 3), (0x05, 1, 3), (0x05, 3, 3), (0x06, 2, 0), (0x06, 3, 0), (0x07, 3, 0), (0x07, 2, 0),
                           (0x01, 2, 0), (0x01, 3, 0), (0x01, 3, 1), (0x01, 2, 1), (0x01, 0, 0x11)):
        emit(itype(opcode, rs, rt, 2))                  # beq beq bne bne blez blez bgtz bgtz bltz bltz bgez bgez bal
        emit(itype(0b001001, 3, 3, 1))                  # delay slot: addiu $3, $3, 1
        emit(itype(0b001001, 2, 2, 1))                  # skipped when taken: addiu $2, $2, 1
    at = ENTRY + 4 * len(p)
    emit(jtype(0x02, (at + 12) >> 2))                   # j over one instruction
    emit(SYNC); emit(itype(0b001001, 2, 2, 1))
    at = ENTRY + 4 * len(p)
    emit(jtype(0x03, (at + 12) >> 2))                   # jal
    emit(SYNC); emit(itype(0b001001, 2, 2, 1))
    at = ENTRY + 4 * len(p)
    emit(itype(0b001101, 0, 30, at + 20))               # ori $30, $0, target
    emit(rtype(0x08, rs=30))                            # jr $30
    emit(SYNC); emit(itype(0b001001, 2, 2, 1)); emit(SYNC)
    at = ENTRY + 4 * len(p)
    emit(itype(0b001101, 0, 30, at + 20))
    emit(rtype(0x09, rs=30, rd=29))                     # jalr $29, $30
    emit(SYNC); emit(itype(0b001001, 2, 2, 1)); emit(SYNC)
    if with_syscalls:
        def li(reg, v):
            if v >> 16:
                emit(itype(0b001111, 0, reg, v >> 16))              # lui
                emit(itype(0b001101, reg, reg, v & 0xFFFF))         # ori
            else:
                emit(itype(0b001101, 0, reg, v))

        def sys(num, a0=0, a1=0, a2=0):
            li(2, num); li(4, a0); li(5, a1); li(6, a2)
            emit(rtype(0b001100))                                   # syscall
        sys(4045, 0x5000)                # brk above the current break
        sys(4045, 0)                     # brk below
        sys(4090, 0, 0x2000)             # mmap, aligned size, from the heap
        sys(4210, 0, 0x1234)             # mmap2, unaligned size
        sys(4090, 0x7000, 0x1000)        # mmap at a given address
        sys(4120)                        # clone
        sys(4003, 0)                     # read stdin
        sys(4003, 5)                     # read bad fd
        sys(4004, 1, 0, 17)              # write stdout
        sys(4004, 7, 0, 17)              # write bad fd
        sys(4055, 0); sys(4055, 2); sys(4055, 9)   # fcntl
        sys(4283, 0x1234)                # set_thread_area
        sys(4999)                        # unknown number
        sys(0x010109, KDATA, KECCAK_LEN, KOUT)      # keccak precompile
        sys(0x00300105, WDATA, 0)        # sha extend
        sys(0x00010106, WDATA, HDATA)    # sha compress
        emit(itype(0b100011, 28, 15, KOUT - DATA))  # lw: read back the first digest word
        if sha_blocks:                   # a "sha2 guest" loop: extend + compress the same block sha_blocks times, chaining hx
            li(16, sha_blocks)
            loop = ENTRY + 4 * len(p)
            sys(0x00300105, WDATA, 0)
            sys(0x00010106, WDATA, HDATA)
            emit(itype(0b001001, 16, 16, 0xFFFF))                   # addiu $16, $16, -1
            at = ENTRY + 4 * len(p)
            emit(itype(0x05, 16, 0, ((loop - (at + 4)) >> 2) & 0xFFFF))   # bne $16, $0, loop
            emit(SYNC)                                              # delay slot
    end = ENTRY + 4 * len(p)
    for _ in range(4):
        emit(SYNC)
    image = {ENTRY + 4 * i: w for i, w in enumerate(p)}
    data = [0x89ABCDEF, 0x01234567, 0x7F80FF00, 0xDEADBEEF, 0, 0x11223344, 0x55667788, 0x99AABBCC, 0xFFFFFFFF, 0xCAFEBABE]
    image[DATA - 4] = 0x0BADF00D
    for i, w in enumerate(data):
        image[DATA + 4 * i] = w
    if with_syscalls:
        x = 0x9E3779B9
        for base, count in ((KDATA, KECCAK_LEN // 4), (WDATA, 16), (HDATA, 8)):
            for i in range(count):
                x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
                image[base + 4 * i] = x
    return image, end
