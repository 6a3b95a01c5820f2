// LUBM vocabulary and the closed-form vertex-id scheme of the synthetic generator.
//
// Id conventions follow the reference converter (datagen/generate_data.cpp:113-125):
//   0 = __PREDICATE__, 1 = rdf:type, index (predicate/type) ids from 2, normal ids from 1<<17.
// LUBM has 17 predicates (incl. rdf:type) and 14 instance classes.
#pragma once
#include <stdint.h>

#define UB "<http://swat.cse.lehigh.edu/onto/univ-bench.owl#"

enum {
    P_PREDICATE = 0,
    P_TYPE = 1,
    // predicates
    P_NAME = 2,
    P_EMAIL = 3,
    P_TELEPHONE = 4,
    P_MEMBEROF = 5,
    P_WORKSFOR = 6,
    P_SUBORG = 7,
    P_UGDEGREE = 8,
    P_MSDEGREE = 9,
    P_DRDEGREE = 10,
    P_ADVISOR = 11,
    P_TAKESCOURSE = 12,
    P_TEACHEROF = 13,
    P_TAOF = 14,
    P_RESEARCHINT = 15,
    P_HEADOF = 16,
    P_PUBAUTHOR = 17,
    // types
    T_UNIVERSITY = 18,
    T_DEPARTMENT = 19,
    T_FULLPROF = 20,
    T_ASSOCPROF = 21,
    T_ASSTPROF = 22,
    T_LECTURER = 23,
    T_UGSTUDENT = 24,
    T_GRADSTUDENT = 25,
    T_COURSE = 26,
    T_GRADCOURSE = 27,
    T_RESEARCHGROUP = 28,
    T_PUBLICATION = 29,
    T_TEACHASSIST = 30,
    T_RESEARCHASSIST = 31,
    LUBM_NUM_INDEX_IDS = 32
};

static inline const char *lubm_index_string(int id) {
    static const char *S[LUBM_NUM_INDEX_IDS] = {
        "__PREDICATE__",
        "<http://www.w3.org/1999/02/22-rdf-syntax-ns#type>",
        UB "name>", UB "emailAddress>", UB "telephone>", UB "memberOf>", UB "worksFor>",
        UB "subOrganizationOf>", UB "undergraduateDegreeFrom>", UB "mastersDegreeFrom>",
        UB "doctoralDegreeFrom>", UB "advisor>", UB "takesCourse>", UB "teacherOf>",
        UB "teachingAssistantOf>", UB "researchInterest>", UB "headOf>", UB "publicationAuthor>",
        UB "University>", UB "Department>", UB "FullProfessor>", UB "AssociateProfessor>",
        UB "AssistantProfessor>", UB "Lecturer>", UB "UndergraduateStudent>", UB "GraduateStudent>",
        UB "Course>", UB "GraduateCourse>", UB "ResearchGroup>", UB "Publication>",
        UB "TeachingAssistant>", UB "ResearchAssistant>"};
    return (id >= 0 && id < LUBM_NUM_INDEX_IDS) ? S[id] : "";
}

// ---- normal vertex ids ---------------------------------------------------------------------
//  [VID_BASE, VID_BASE + 2^17)            shared literal pool (telephone, names, research interests)
//  [UNIV_BASE + u*(2^17-1), +2^17-1)      block of university u.  The block length is odd (a Mersenne prime) on purpose: with a
//                                         power of two every university -- and every d-th department -- would land on the same
//                                         shard of a vid % n cluster, which ids handed out in order of first appearance (the
//                                         reference's id mapping of real UBA output) never do
//      local 0       the university       local 1  its name literal
//      local 2+d     department d (d<25)  local 64.. everything else, in generation order
#define LUBM_VID_BASE (1u << 17)
#define LUBM_UNIV_BLOCK ((1u << 17) - 1u)
#define LUBM_UNIV_BASE (LUBM_VID_BASE + (1u << 17))
#define LUBM_LOCAL_UNIV_NAME 1u
#define LUBM_LOCAL_DEPT0 2u
#define LUBM_LOCAL_FIRST_FREE 64u

#define LUBM_LIT_TELEPHONE (LUBM_VID_BASE + 0u)
#define LUBM_NUM_RESEARCH 30u

enum {
    NP_DEPT = 0, NP_FULLPROF, NP_ASSOCPROF, NP_ASSTPROF, NP_LECTURER, NP_UGSTUDENT,
    NP_GRADSTUDENT, NP_COURSE, NP_GRADCOURSE, NP_RESEARCHGROUP, NP_PUBLICATION, NP_COUNT
};

static inline uint32_t lubm_univ_id(uint32_t u) { return LUBM_UNIV_BASE + u * LUBM_UNIV_BLOCK; }
static inline uint32_t lubm_dept_id(uint32_t u, uint32_t d) { return lubm_univ_id(u) + LUBM_LOCAL_DEPT0 + d; }
static inline uint32_t lubm_research_literal(uint32_t i) { return LUBM_VID_BASE + 16u + i; }
static inline uint32_t lubm_name_literal(int pool, uint32_t idx) {
    return LUBM_VID_BASE + 1024u + (uint32_t)pool * 4096u + idx;
}
// largest scale whose ids fit in 32 bits
#define LUBM_MAX_UNIVS ((0xFFFFFFFFu - LUBM_UNIV_BASE) / LUBM_UNIV_BLOCK)
